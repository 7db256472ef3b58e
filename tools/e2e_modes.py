"""Compare the two wf_step_host transports on the GPU box: staged cudaMemcpyAsync ("copy") vs
zero-copy stores into page-locked host memory ("direct").  Usage: python tools/e2e_modes.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wildfire_control_python_b200 import BatchedForestFire  # noqa: E402

for mode in ("copy", "direct", "hybrid"):
    os.environ["WF_HOST_MODE"] = mode
    for N in (4096, 8192, 65536):
        env = BatchedForestFire(N, width=14, height=14, auto_reset=True, seed=0)
        env.reset()
        rng = np.random.default_rng(0)
        acts = rng.integers(0, 4, size=(103, N), dtype=np.int32)
        for k in range(3):
            env.step_host(acts[k])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(100):
            env.step_host(acts[3 + k])
        dt = time.perf_counter() - t0
        print(mode, N, f"{dt / 100 * 1e6:.1f} us/step  {N * 100 / dt:.3e} env-steps/s", flush=True)
        env.close()
