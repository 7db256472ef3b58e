"""A/B helper: us/step of wf_rollout on one workload with the library named by WILDFIRE_B200_LIB.
    WILDFIRE_B200_LIB=build_ab/libold.so python tools/ab_rollout.py c4|c5|c2 [steps_per_launch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS  # noqa: E402
from wildfire_control_python_b200 import BatchedForestFire  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
wl = WORKLOADS[name]
N, K = int(os.environ.get("AB_N", wl["n_envs"])), (int(sys.argv[2]) if len(sys.argv) > 2 else wl["chunk"])
W, H = wl["meta"]["width"], wl["meta"]["height"]
env = BatchedForestFire(N, auto_reset=True, seed=0, **wl["meta"])
env.reset()
out = (torch.empty((K, N, W, H, 3), dtype=torch.uint8, device="cuda"), torch.empty((K, N), dtype=torch.float64, device="cuda"),
       torch.empty((K, N), dtype=torch.uint8, device="cuda"))
for _ in range(4):
    env.rollout(K, out=out)
res = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    n = max(2, (20 if name != "c2" else 100) * wl["chunk"] // K)
    for _ in range(n):
        env.rollout(K, out=out)
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) * 1e3 / (n * K))
print(name, f"K={K}", os.environ.get("WILDFIRE_B200_LIB", "in-tree"), " ".join(f"{r:.2f}" for r in res), "us/step", flush=True)
