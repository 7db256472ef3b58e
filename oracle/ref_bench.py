"""Time the UNMODIFIED Python reference (Simulation/forest_fire.py: ForestFire.reset / step) on host cores.

TEST / MEASUREMENT INFRASTRUCTURE -- used only by bench.py's `cpu_baseline.python_ref` leg and `--impl reference`.
SURVEY.md 8(d) "CPU baseline timing": the reference ``ForestFire`` in P worker processes (multiprocessing), one batch
of envs per process, ``time.perf_counter`` around the step loop only, env-steps summed.  The reference runs through
its own public API and stock code path: its own RNG (no Philox patching is needed to TIME it), its own A* (compiled by
`make -C oracle ref` from pyastar/astar.cpp where it lies).  On the GPU box the package is imported from the
git-ignored snapshot oracle/_ref/ (see oracle/Makefile), in the build container from /root/reference.
"""
from __future__ import annotations

import json
import subprocess
import sys
import os
import time

import numpy as np


def available() -> bool:
    from . import ref_harness
    return ref_harness.reference_available() and os.path.isfile(os.path.join(ref_harness.REF_BUILD, "pyastar", "astar.so"))


def _worker(rank: int, meta: dict, n_envs: int, warmup: int, steps: int) -> None:
    """One process: build the envs, say READY, wait for GO on stdin, run, print the result as JSON."""
    from . import ref_harness
    constants, utility, environment, forest_fire = ref_harness.load_reference()
    m = constants.METADATA
    W, H = int(meta["width"]), int(meta["height"])
    m["width"], m["height"] = W, H
    wind = meta.get("wind", [0.54, (0, 0)])
    m["wind"] = wind if wind == "random" else [wind[0], tuple(wind[1])]
    for k in ("a_speed", "n_actions", "make_rivers", "allow_dig_toggle"):
        if k in meta:
            m[k] = meta[k]
    m["a_speed_iter"] = m["a_speed"]
    environment.WIDTH, environment.HEIGHT = W, H  # frozen at import (environment.py:22-23): re-point before World()
    rng = np.random.default_rng(1000 + rank)
    sims = [forest_fire.ForestFire() for _ in range(n_envs)]
    extra = int(meta.get("extra_ignitions", 0))

    def reset(sim):
        sim.reset()
        for _ in range(extra):  # the workload's extra ignitions, through the public World.set_fire_to
            sim.W.set_fire_to((int(rng.integers(0, W)), int(rng.integers(0, H))))

    for sim in sims:
        reset(sim)
    actions = rng.integers(0, int(m["n_actions"]), size=(warmup + steps, n_envs))

    def run_steps(k0, k1):
        n = 0
        for k in range(k0, k1):
            row = actions[k]
            for i, sim in enumerate(sims):
                done = sim.step(int(row[i]))[2]
                n += 1
                if done:
                    reset(sim)
        return n

    print("READY", flush=True)
    sys.stdin.readline()  # GO
    run_steps(0, warmup)
    t0 = time.perf_counter()
    n = run_steps(warmup, warmup + steps)
    t1 = time.perf_counter()
    print(json.dumps({"rank": rank, "n": n, "t0": t0, "t1": t1}), flush=True)


def run(meta: dict, n_envs_total: int, steps: int, warmup: int, procs: int = 0):
    """``steps`` timed steps over ``n_envs_total`` reference envs spread over ``procs`` processes (default: all cores).
    Returns dict(value=env-steps/s, env_steps, seconds, procs, envs).  Separate interpreters (the parent may hold a
    CUDA context: never fork it), started together once every one has built its envs."""
    procs = procs or (os.cpu_count() or 1)
    procs = max(1, min(procs, n_envs_total))
    per = [n_envs_total // procs + (1 if r < n_envs_total % procs else 0) for r in range(procs)]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    meta_json = json.dumps({k: (list(v) if isinstance(v, tuple) else v) for k, v in meta.items()})
    ps = [subprocess.Popen([sys.executable, "-m", "oracle.ref_bench", "--worker", str(r), meta_json, str(per[r]), str(warmup), str(steps)],
                           cwd=root, stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True) for r in range(procs)]
    try:
        for p in ps:
            line = p.stdout.readline()
            if line.strip() != "READY":
                raise RuntimeError(f"reference worker failed to start: {line!r}")
        for p in ps:
            p.stdin.write("GO\n")
            p.stdin.flush()
        res = [json.loads(p.stdout.readline()) for p in ps]
    finally:
        for p in ps:
            try:
                p.stdin.close()
                p.wait(timeout=30)
            except Exception:
                p.kill()
    n = sum(r["n"] for r in res)
    t0, t1 = min(r["t0"] for r in res), max(r["t1"] for r in res)  # CLOCK_MONOTONIC is system-wide: the slowest worker counts
    return dict(value=n / (t1 - t0), env_steps=n, seconds=t1 - t0, procs=procs, envs=n_envs_total)


if __name__ == "__main__":
    if len(sys.argv) >= 7 and sys.argv[1] == "--worker":
        _worker(int(sys.argv[2]), json.loads(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]))
    else:
        print(json.dumps(run(dict(width=14, height=14), 4 * (os.cpu_count() or 1), 50, 5)))
