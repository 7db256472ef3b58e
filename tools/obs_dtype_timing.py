"""Time wf_rollout of the tile family (256x256) with uint8 / bfloat16 / float32 observations.
    python tools/obs_dtype_timing.py [n_envs] [K]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wildfire_control_python_b200 import BatchedForestFire  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
for dt in (torch.uint8, torch.bfloat16, torch.float32):
    env = BatchedForestFire(N, obs_dtype=dt, auto_reset=True, seed=0, width=256, height=256, wind=[0.85, (1, 0)], extra_ignitions=32)
    env.reset()
    out = (torch.empty((K, N, 256, 256, 3), dtype=dt, device="cuda"), torch.empty((K, N), dtype=torch.float64, device="cuda"),
           torch.empty((K, N), dtype=torch.uint8, device="cuda"))
    for _ in range(3):
        env.rollout(K, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    reps = 10
    for _ in range(reps):
        env.rollout(K, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * K)
    gb = out[0].element_size() * N * 256 * 256 * 3 / 1e9
    print(f"{str(dt):16s} {us:8.1f} us/step  obs write {gb / (us * 1e-6):7.0f} GB/s", flush=True)
    env.close()
    del out, env
    torch.cuda.empty_cache()
