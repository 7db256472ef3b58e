"""`main.py -r -t <TYPE> -m M -e E` of the reference, against the CUDA environment and the torch learners.

    python tools/train_demo.py BOTH 100 600 [size]
"""
import sys
import time

import numpy as np

from wildfire_control_python_b200 import ForestFire
from wildfire_control_python_b200 import agents as A

kind = sys.argv[1] if len(sys.argv) > 1 else "BOTH"
memories = int(sys.argv[2]) if len(sys.argv) > 2 else 100
episodes = int(sys.argv[3]) if len(sys.argv) > 3 else 600
size = int(sys.argv[4]) if len(sys.argv) > 4 else 10
cls = {"DQN": A.DQN, "SARSA": A.DQN_SARSA, "DDQN": A.DQN_DUEL, "BOTH": A.DQN_BOTH}[kind]
sim = ForestFire(width=size, height=size, seed=1)
agent = cls(sim, name=f"demo_{kind}", verbose=False)
agent.out_dir = "gpurun_out"
t0 = time.time()
steps = agent.collect_memories_batched(memories, n_envs=1024, k_steps=256, seed=7)
print(f"{len(agent.memory)} demonstration transitions from {memories} contained episodes "
      f"({steps} env-steps simulated) in {time.time() - t0:.2f} s", flush=True)
t0 = time.time()
agent.learn(episodes)
r = np.array(agent.logs["total_rewards"])
print(f"{kind}: {episodes} episodes in {time.time() - t0:.1f} s; mean return first 100: {r[:100].mean():.0f}, "
      f"last 100: {r[-100:].mean():.0f}; deaths last 100: {sum(agent.logs['agent_deaths'][-100:])}", flush=True)
