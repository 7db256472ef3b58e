"""A/B of every wf_step_host transport on one GPU box (C2 shape: N envs of 14x14), each in its own subprocess because the
WF_HOST_* switches are read once per handle / process:

    launch      default launch-per-step path (kernel -> HBM staging -> one DMA copy -> host-thread expansion)
    graph       the same with kernel + copy as one CUDA graph (WF_HOST_GRAPH=1)
    direct      the kernel stores the bit stream straight into mapped host memory (WF_HOST_PACKED=direct)
    u8          the uint8 array over PCIe, no host threads (WF_HOST_PACKED=0)
    session     wf_host_session: resident step-server kernel, doorbell + per-slice completion flags
    session/T   ... with T host threads;  noavx512: the AVX2 expansion instead of AVX-512

Usage: python tools/e2e_ab.py [n_envs] [steps]      (prints us/step, best and median of 5 blocks)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = r"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, %r)
import torch
from wildfire_control_python_b200 import BatchedForestFire
N, K, session = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3] != "0"
env = BatchedForestFire(N, width=14, height=14, auto_reset=True, seed=0)
env.reset()
if session:
    assert env.host_session(True, persistent_obs=sys.argv[3] == "2")
acts = np.random.default_rng(0).integers(0, 4, size=(256, N), dtype=np.int32)
for k in range(50):
    env.step_host(acts[k])
blocks = []
for b in range(5):
    t0 = time.perf_counter()
    for k in range(K):
        env.step_host(acts[k & 255])
    blocks.append((time.perf_counter() - t0) / K * 1e6)
print(json.dumps({"best": min(blocks), "median": sorted(blocks)[2], "threads": env.host_threads, "session": env.host_session_state}))
env.close()
""" % ROOT


def run(name, env_extra, session, N, K):
    env = dict(os.environ, WF_HOST_TIMING="1", **env_extra)
    r = subprocess.run([sys.executable, "-c", WORKER, str(N), str(K), str(int(session))], env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(f"{name:22s} FAILED: {r.stderr[-400:]}", flush=True)
        return
    d = json.loads(line[-1])
    extra = " | ".join(l for l in r.stderr.splitlines() if l.startswith("wf_"))
    if os.environ.get("E2E_AB_FULL"):
        print(r.stderr, flush=True)
    print(f"{name:22s} best {d['best']:6.1f}  median {d['median']:6.1f} us/step   threads {d['threads']:2d}  {extra}", flush=True)


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    run("launch", {}, False, N, K)
    if len(sys.argv) > 3 and sys.argv[3] == "quick":
        for t in (12, 15, 6, 3):
            run(f"session/{t}", {"WF_HOST_THREADS": str(t)}, True, N, K)
        for t in (12, 15, 6, 3):
            run(f"session/{t} flag+fence", {"WF_HOST_THREADS": str(t), "WF_SESSION_SECTORS": "0"}, True, N, K)
        sys.exit(0)
    if len(sys.argv) > 3 and sys.argv[3] == "persistent":
        for t in (12, 6, 4, 3, 2, 1):
            run(f"session/{t}", {"WF_HOST_THREADS": str(t)}, 1, N, K)
            run(f"persistent/{t}", {"WF_HOST_THREADS": str(t)}, 2, N, K)
        sys.exit(0)
    run("graph", {"WF_HOST_GRAPH": "1"}, False, N, K)
    run("direct", {"WF_HOST_PACKED": "direct"}, False, N, K)
    run("u8", {"WF_HOST_PACKED": "0"}, False, N, K)
    run("session", {}, True, N, K)
    for t in (1, 2, 3, 4, 6, 8, 12, 15):
        run(f"session/{t}", {"WF_HOST_THREADS": str(t)}, True, N, K)
    run("session noavx512", {"WF_HOST_NO_AVX512": "1"}, True, N, K)
    run("session/3 noavx512", {"WF_HOST_NO_AVX512": "1", "WF_HOST_THREADS": "3"}, True, N, K)
    run("launch noavx512", {"WF_HOST_NO_AVX512": "1"}, False, N, K)
