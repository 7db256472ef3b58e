"""Mean return of the reference's trained networks (tests/golden/keras) on the CUDA env vs the reference's logs."""
import json
import os
import time

import numpy as np

from wildfire_control_python_b200 import ForestFire
from wildfire_control_python_b200 import agents as A

HERE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "keras")
KAT = json.load(open(os.path.join(HERE, "kat.json")))
for name, meta in sorted(KAT.items()):
    cls = A.DQN_BOTH if name.startswith("BOTH") else A.DQN_SARSA
    ag = cls(ForestFire(width=meta["size"], height=meta["size"], seed=1), verbose=False)
    ag.load_keras_weights(os.path.join(HERE, name))
    t0 = time.time()
    rets, died = ag.evaluate_batched(n_envs=8192, episodes_per_env=4, eps=meta["log"]["min_eps"], seed=3)
    print(f"{name}: {len(rets)} episodes in {time.time() - t0:.1f} s: mean return {rets.mean():.1f} +- {rets.std() / np.sqrt(len(rets)):.1f}, "
          f"deaths {died.mean():.4f} | reference log (last 2500 training episodes): {meta['log']['mean_last_2500']:.1f}, "
          f"deaths {meta['log']['death_rate_last_2500']:.4f}", flush=True)
