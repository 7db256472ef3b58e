"""bench.py's JSON line, assembled on the CPU from canned measurements: every key the bench contract names must be there
and consistent (value = envs x steps / time, roofline.frac = achieved / peak, ...).  The GPU legs are stubbed; the
cpu_baseline leg runs the real C oracle for a fraction of a second."""
import json
import sys
import types

import pytest

import bench


class _FakeDist:
    rank, world, local = 0, 1, 0

    def __init__(self, gpus):
        pass

    def close(self):
        pass


class _NoClocks:
    def __init__(self, index):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def summary(self):
        return {"sm_mhz": 1965, "sm_max_mhz": 1965, "reasons": [], "samples": 3}


def _canned(name, K):
    wl = bench.WORKLOADS[name]
    N, W, H = wl["n_envs"], wl["meta"]["width"], wl["meta"]["height"]
    chunk = min(wl["chunk"], K)
    ms = 0.0025 * K if W <= 32 else 0.05 * K
    fam = "warp" if W <= 32 else "tile"
    return {"name": name, "wl": wl, "N": N, "W": W, "H": H, "K": K, "chunk": chunk, "fused": True, "graph": False, "ms": ms,
            "launches": -(-K // chunk), "value": N * K / (ms * 1e-3), "obs_mb": chunk * N * W * H * 3 / 1e6, "family": fam,
            "state_bytes": 1280 if fam == "warp" else 409664,
            "stats": {"env_steps": N * K, "episodes": 1, "deaths": 1, "contained": 0, "burnouts": 0, "ticks": 0},
            "e2e": {"value": 9e7, "unit": "env-steps/s", "h2d_bytes_per_step": N * 4, "d2h_bytes_per_step": 1, "steps": 3},
            "per_step": {"value": 3e8, "unit": "env-steps/s", "steps": 3}}


@pytest.mark.parametrize("argv", [[], ["--workload", "c4", "--steps", "64", "--warmup", "8"], ["--steps", "10", "--warmup", "3"],
                                  ["--only-value", "--steps", "512"]], ids=["default", "c4", "short", "only_value"])
def test_bench_line_has_the_contract_keys(monkeypatch, capsys, argv):
    monkeypatch.setattr(bench, "Dist", _FakeDist)
    monkeypatch.setattr(bench, "ClockSampler", _NoClocks)
    monkeypatch.setattr(bench, "CPU_BASELINE_SECONDS", 0.2)
    monkeypatch.setattr(bench, "measure", lambda D, name, K, Wm, chunk, no_graph, e2e, per_step: {
        k: v for k, v in _canned(name, K).items() if (k != "e2e" or e2e) and (k != "per_step" or per_step)})
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert key in line, key
    assert line["metric"] == "env_steps_per_sec" and line["unit"] == "env-steps/s" and line["scaling"] == "weak"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["n_gpus"] == 1
    assert "workload" in line["config"] and "model" not in line["config"] and "l2" in line["config"]
    assert line["warmup"] >= 3 and line["gpu_launches"] >= 1
    wl = bench.WORKLOADS[argv[argv.index("--workload") + 1] if "--workload" in argv else "c2"]
    assert line["value"] == pytest.approx(wl["n_envs"] * line["steps"] / (line["ms_per_step"] * line["steps"] * 1e-3))
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"])
    assert r["bytes_per_unit"] == wl["meta"]["width"] * wl["meta"]["height"] * 15
    assert r["achieved"] == pytest.approx(line["value"] * r["bytes_per_unit"] / 1e9)
    assert r["units_per_launch"] == wl["n_envs"] * line["config"]["steps_per_launch"]
    if "--only-value" in argv:
        assert line["e2e"] is None and line["cpu_baseline"] is None and "secondary" not in line
    else:
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
        cb = line["cpu_baseline"]
        assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] > 0 and "sample" in cb
        assert ("secondary" in line) == ("--workload" not in argv)


def test_reference_arm_line(monkeypatch, capsys):
    """--impl reference: the oracle port on the host cores, same metric / unit / config keys, impl = reference."""
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--steps", "3", "--warmup", "1"])
    monkeypatch.setitem(bench.WORKLOADS, "c2", dict(bench.WORKLOADS["c2"], n_envs=64))
    bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "env_steps_per_sec" and line["unit"] == "env-steps/s"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["higher_is_better"] is True
