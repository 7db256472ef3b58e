"""Golden inputs for the Keras-weights reader and the trained-policy parity test (TEST INFRASTRUCTURE).

Copies two of the reference's trained networks (Models/<size>/<name>, written by DQN.save_model,
DQN.py:441-443) into tests/golden/keras/ and records, next to them, what the reader must return
(shape and SHA-256 of every weight array) and what the reference's own training log says about the
policy (Logs/<size>/<name>: mean total reward and death rate of the last 2500 / 500 episodes, the
quantity the thesis tabulates).  Run in the build container, where /root/reference exists:

    python oracle/gen_keras_fixture.py
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wildfire_control_python_b200.keras_h5 import read_keras_weights  # noqa: E402

REF = "/root/reference"
PICKS = [("10-sized", "BOTH9-10s-10k-36761m-06-20-0555"), ("14-sized", "SARSA9-14s-10k-47298m-06-24-0808")]
OUT = os.path.join(ROOT, "tests", "golden", "keras")

kat = {}
for size, name in PICKS:
    src = os.path.join(REF, "Models", size, name)
    shutil.copyfile(src, os.path.join(OUT, name))
    os.chmod(os.path.join(OUT, name), 0o644)
    w = read_keras_weights(src)
    log = json.load(open(os.path.join(REF, "Logs", size, name)))
    tr = np.array(log["total_rewards"], dtype=np.float64)
    deaths = np.array(log["agent_deaths"], dtype=np.float64)
    kat[name] = {
        "size": int(log["metadata"]["width"]),
        "file_sha256": hashlib.sha256(open(src, "rb").read()).hexdigest(),
        "arrays": {k: {"shape": list(v.shape), "dtype": str(v.dtype), "sha256": hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest(),
                       "first": float(v.ravel()[0]), "sum": float(v.astype(np.float64).sum())} for k, v in sorted(w.items())},
        "log": {"episodes": len(tr), "mean_last_2500": float(tr[-2500:].mean()), "mean_last_500": float(tr[-500:].mean()),
                "death_rate_last_2500": float(deaths[-2500:].mean()), "min_eps": log["metadata"]["min_eps"],
                "metadata": {k: log["metadata"][k] for k in ("width", "height", "wind", "a_speed", "n_actions", "make_rivers",
                                                             "contained_bonus", "death_penalty", "default_reward")}},
    }
json.dump(kat, open(os.path.join(OUT, "kat.json"), "w"), indent=1)
# bench.py --workload c3 rolls the 14-sized network out as its policy: the four arrays as a plain .npz (a benchmark INPUT,
# kept outside tests/)
w = read_keras_weights(os.path.join(OUT, PICKS[1][1]))
os.makedirs(os.path.join(ROOT, "bench_inputs"), exist_ok=True)
np.savez(os.path.join(ROOT, "bench_inputs", "c3_policy_sarsa9_14s.npz"), kernel1=w["dense_1/kernel:0"], bias1=w["dense_1/bias:0"],
         kernel2=w["dense_2/kernel:0"], bias2=w["dense_2/bias:0"])
print(json.dumps({k: v["log"] for k, v in kat.items()}, indent=1))
