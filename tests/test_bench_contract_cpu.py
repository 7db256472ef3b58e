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
    chunk = wl["chunk"]
    us = 2.5 if W <= 32 else 50.0
    R = bench.plan_repeats(K, chunk, us)
    ms = us * 1e-3 * R * K
    fam = "warp" if W <= 32 else "tile"
    out = {"name": name, "wl": wl, "N": N, "W": W, "H": H, "K": K, "repeats": R, "chunk": chunk, "ms": ms,
           "launches": R * K // chunk, "launches_total": R * K // chunk + 4, "value": N * R * K / (ms * 1e-3),
           "obs_mb": chunk * N * W * H * 3 / 1e6, "family": fam, "state_bytes": 1280 if fam == "warp" else 409664,
           "tile_geometry": [0, 0] if fam == "warp" else [256, 1],
           "stats": {"env_steps": N * K, "episodes": 1, "deaths": 1, "contained": 0, "burnouts": 0, "ticks": N * K},
           "e2e": {"value": 9e7, "unit": "env-steps/s", "h2d_bytes_per_step": N * 4, "d2h_bytes_per_step": 1, "steps": K, "repeats": 3},
           "per_step": {"value": 3e8, "unit": "env-steps/s", "steps": 3}}
    if wl.get("policy"):
        out["via_torch"] = {"value": 1e8, "unit": "env-steps/s"}
        out["policy"] = "canned"
    return out


def test_plan_repeats_fills_the_region_with_whole_launches():
    for K, chunk, us in [(20, 256, 2.3), (20, 16, 75.0), (2048, 256, 2.3), (7, 64, 10.0), (20000, 256, 2.3), (1, 256, 2.0)]:
        R = bench.plan_repeats(K, chunk, us, 0.3)
        assert (R * K) % chunk == 0 and R * K * us * 1e-6 >= 0.3
        assert (R - chunk) * K * us * 1e-6 < 0.3 or R <= chunk  # not more than one rounding unit too long
    assert bench.plan_repeats(20, 256, 2.3, 0.0) == 64  # the fewest whole launches (profiling runs)


@pytest.mark.parametrize("argv", [[], ["--workload", "c4", "--steps", "64", "--warmup", "8"], ["--steps", "20", "--warmup", "5"],
                                  ["--only-value", "--steps", "512"], ["--workload", "c3", "--steps", "20"]],
                         ids=["default", "c4", "driver_20_5", "only_value", "c3"])
def test_bench_line_has_the_contract_keys(monkeypatch, capsys, argv):
    monkeypatch.setattr(bench, "Dist", _FakeDist)
    monkeypatch.setattr(bench, "ClockSampler", _NoClocks)
    monkeypatch.setattr(bench, "CPU_BASELINE_SECONDS", 0.2)
    monkeypatch.setattr(bench, "python_ref_baseline", lambda wl, steps, warmup, budget_s=60.0: None)
    monkeypatch.setattr(bench, "measure", lambda D, name, K, Wm, chunk, e2e, per_step: {
        k: v for k, v in _canned(name, K).items() if (k != "e2e" or e2e) and (k != "per_step" or per_step)})
    monkeypatch.setattr(bench, "measure_policy", lambda D, name, K, Wm: {k: v for k, v in _canned(name, K).items() if k not in ("e2e", "per_step")})
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "repeats", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "launch", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert key in line, key
    assert line["metric"] == "env_steps_per_sec" and line["unit"] == "env-steps/s" and line["scaling"] == "weak"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["n_gpus"] == 1
    name = argv[argv.index("--workload") + 1] if "--workload" in argv else "c2"
    wl = bench.WORKLOADS[name]
    assert "model" not in line["config"] and "l2" in line["config"]
    assert {k: line["config"][k] for k in ("workload", "grid", "envs_per_gpu")} == bench.workload_config(name)
    assert line["warmup"] >= 3 and line["gpu_launches"] >= 1
    assert line["steps"] == (int(argv[argv.index("--steps") + 1]) if "--steps" in argv else 2048)
    assert line["value"] == pytest.approx(wl["n_envs"] / (line["ms_per_step"] * 1e-3))
    assert line["ms_per_step"] * line["steps"] * line["repeats"] >= 300.0  # the timed region lasted >= 0.3 s
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"])
    assert r["achieved"] == pytest.approx(line["value"] * r["bytes_per_unit"] / 1e9)
    assert r["units_per_launch"] == wl["n_envs"] * line["launch"]["steps_per_launch"]
    sa = r["survey_accounting"]
    assert sa["bytes_per_unit"] == wl["meta"]["width"] * wl["meta"]["height"] * 15
    assert sa["frac"] == pytest.approx(line["value"] * sa["bytes_per_unit"] / 1e9 / r["peak"])
    if r["traffic"] is not None:
        assert r["traffic_over_algorithmic"] == pytest.approx(r["traffic"] / (r["bytes_per_unit"] * r["units_per_launch"]))
    if name in ("c2", "c4", "c5"):
        assert r["traffic"] is not None  # profiles/ncu_traffic.json covers every non-policy workload
    if "--only-value" in argv or name == "c3":
        assert line["e2e"] is None and "secondary" not in line
    else:
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
        cb = line["cpu_baseline"]
        assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] > 0 and "sample" in cb and "python_ref" in cb
        assert ("secondary" in line) == ("--workload" not in argv)
        if "secondary" in line:
            assert set(line["secondary"]) == {"c4", "c5", "c3"}
            for k, e in line["secondary"].items():
                assert e["config"] == bench.workload_config(k) and e["roofline"]["frac"] > 0 and e["value"] > 0
            assert "via_torch" in line["secondary"]["c3"]


def test_reference_arm_line(monkeypatch, capsys):
    """--impl reference: the reference's Python step (when the snapshot / tree is there) and the oracle port on the host
    cores; same metric / unit / config as our arm, impl = reference."""
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--steps", "3", "--warmup", "1"])
    monkeypatch.setitem(bench.WORKLOADS, "c2", dict(bench.WORKLOADS["c2"], n_envs=64))
    bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "env_steps_per_sec" and line["unit"] == "env-steps/s"
    assert line["config"] == bench.workload_config("c2")
    cb = line["cpu_baseline"]
    assert line["value"] > 0 and cb["value"] == line["value"]
    from oracle import ref_bench
    if ref_bench.available():
        assert cb["kind"] == "_ref" and cb["port"]["kind"] == "port" and cb["port"]["value"] > cb["value"]
    else:
        assert cb["kind"] == "port"
    assert line["e2e"] == {"value": line["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["higher_is_better"] is True
