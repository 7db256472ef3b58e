import sys
import torch
from wildfire_control_python_b200 import BatchedForestFire
W, H, N, K, pol, ar, ob = sys.argv[1:8]
env = BatchedForestFire(int(N), width=int(W), height=int(H), seed=3, auto_reset=bool(int(ar)))
env.reset()
torch.cuda.synchronize()
o, r, d = env.rollout(int(K), policy=pol, obs=bool(int(ob)))
torch.cuda.synchronize()
print("ok", sys.argv[1:], float(r.sum()), int(d.sum()), flush=True)
