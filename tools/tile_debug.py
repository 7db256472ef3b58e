"""Small tile-family run for compute-sanitizer: reset + rollout on a few grid shapes."""
import sys
import torch
from wildfire_control_python_b200 import BatchedForestFire

shapes = [(40, 36, 41), (64, 64, 8), (256, 256, 4)] if len(sys.argv) < 2 else [tuple(int(v) for v in sys.argv[1].split("x"))]
for W, H, N in shapes:
    env = BatchedForestFire(N, width=W, height=H, seed=3, auto_reset=True)
    env.reset()
    torch.cuda.synchronize()
    print("reset ok", W, H, N, flush=True)
    o, r, d = env.rollout(30, policy="walk")
    torch.cuda.synchronize()
    print("walk ok", float(r.sum()), int(d.sum()), flush=True)
    o, r, d = env.rollout(30)
    torch.cuda.synchronize()
    print("stream ok", float(r.sum()), int(d.sum()), env.stats(), flush=True)
    env.close()
