"""Experiment (next GPU session): wf_step_host's packed path with the kernel + DMA copy as one CUDA graph
(WF_HOST_GRAPH=1, off by default) against the default two API calls.  Checks that both deliver identical
observations / rewards / dones, then times them alternately.  Usage: python tools/e2e_graph_experiment.py [n_envs]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wildfire_control_python_b200 import BatchedForestFire  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
os.environ["WF_HOST_TIMING"] = "1"
rng = np.random.default_rng(0)
acts = rng.integers(0, 4, size=(603, N), dtype=np.int32)


def make(graph):
    os.environ["WF_HOST_GRAPH"] = "1" if graph else "0"  # read when the handle's host path is first used
    env = BatchedForestFire(N, width=14, height=14, auto_reset=True, seed=0)
    env.reset()
    env.step_host(acts[0])
    return env


a, b = make(False), make(True)
for k in range(1, 200):  # identical trajectories, step by step
    oa, ra, da, _ = a.step_host(acts[k])
    ob, rb, db, _ = b.step_host(acts[k])
    assert np.array_equal(oa, ob) and np.array_equal(ra, rb) and np.array_equal(da, db), k
print("graph path == default path on 199 steps", flush=True)
for rep in range(3):
    for name, env in (("default", a), ("graph", b)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(200, 600):
            env.step_host(acts[k])
        dt = time.perf_counter() - t0
        print(f"{name:8s} {dt / 400 * 1e6:6.1f} us/step  {N * 400 / dt:.3e} env-steps/s", flush=True)
a.close()
b.close()
