import torch
from bench import WORKLOADS
from wildfire_control_python_b200 import BatchedForestFire
wl = WORKLOADS["c4"]
N = wl["n_envs"]
env = BatchedForestFire(N, auto_reset=True, seed=0, **wl["meta"])
env.reset()
K = 16
rew = torch.empty((K, N), dtype=torch.float64, device="cuda")
done = torch.empty((K, N), dtype=torch.uint8, device="cuda")
for _ in range(8):
    env.rollout(K, out=(None, rew, done), obs=False)
torch.cuda.synchronize()
print("ok")
