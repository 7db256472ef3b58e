"""Time the tile-family rollout with and without the observation block (c4 / c5 workloads)."""
import sys
import torch
from bench import WORKLOADS
from wildfire_control_python_b200 import BatchedForestFire

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
wl = WORKLOADS[name]
N, W, H = wl["n_envs"], wl["meta"]["width"], wl["meta"]["height"]
env = BatchedForestFire(N, auto_reset=True, seed=0, **wl["meta"])
env.reset()
K = 16
obs = torch.empty((K, N, W, H, 3), dtype=torch.uint8, device="cuda")
rew = torch.empty((K, N), dtype=torch.float64, device="cuda")
done = torch.empty((K, N), dtype=torch.uint8, device="cuda")
for with_obs in (True, False, True, False):
    out = (obs if with_obs else None, rew, done)
    for _ in range(3):
        env.rollout(K, out=out, obs=with_obs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        env.rollout(K, out=out, obs=with_obs)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name} obs={with_obs}: {e0.elapsed_time(e1) * 1e3 / (10 * K):.2f} us/step", flush=True)
print(env.stats())
