"""Every trained network of the reference tree (Models/*/*) rolled out on the C ORACLE (CPU, build container
only: needs /root/reference) and compared with what the reference's own log says about that run.

    python tools/eval_all_reference_models.py [episodes_per_model]

Prints one line per model and a summary for the converged runs (death rate of the last 2500 training
episodes below 5 %), where the final weights are representative of the logged average."""
import glob
import json
import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wf_oracle as wo  # noqa: E402
from wildfire_control_python_b200.keras_h5 import read_keras_weights  # noqa: E402

REF = "/root/reference"


def evaluate(args):
    path, n = args
    size = 10 if "/10-sized/" in path else 14
    name = os.path.basename(path)
    log_path = os.path.join(REF, "Logs", f"{size}-sized", name)
    if not os.path.isfile(log_path) or open(path, "rb").read(8) != b"\x89HDF\r\n\x1a\n":
        return None
    log = json.load(open(log_path))
    tr, deaths = np.array(log["total_rewards"], float), np.array(log["agent_deaths"], float)
    w = {k: v.astype(np.float64) for k, v in read_keras_weights(path).items()}

    def q(obs):
        x = obs.reshape(-1).astype(np.float64)
        h = 1.0 / (1.0 + np.exp(-np.clip(x @ w["dense_1/kernel:0"] + w["dense_1/bias:0"], -60, 60)))
        return h @ w["dense_2/kernel:0"] + w["dense_2/bias:0"]  # the advantage stream decides the argmax

    env = wo.OracleEnv(dict(width=size, height=size, seed=99))
    rng = np.random.default_rng(1)
    rets, died = [], 0
    for _ in range(n):
        o, done, tot = env.reset(), False, 0.0
        while not done:
            a = int(np.argmax(q(o))) if rng.random() > 0.01 else int(rng.integers(4))
            o, r, done, _ = env.step(a)
            tot += r
        rets.append(tot)
        died += int(r == -1000)
    rets = np.array(rets)
    return dict(name=name, size=size, ours=float(rets.mean()), se=float(rets.std() / np.sqrt(n)), ours_deaths=died / n,
                log2500=float(tr[-2500:].mean()), log500=float(tr[-500:].mean()), log_deaths=float(deaths[-2500:].mean()))


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    files = sorted(glob.glob(os.path.join(REF, "Models", "*", "*")))
    with Pool(os.cpu_count()) as pool:
        rows = [r for r in pool.map(evaluate, [(f, n) for f in files]) if r]
    for r in rows:
        print(f"{r['name']:42s} ours {r['ours']:8.1f} +- {r['se']:5.1f} (deaths {r['ours_deaths']:.3f}) | log last 2500: {r['log2500']:8.1f}, "
              f"last 500: {r['log500']:8.1f} (deaths {r['log_deaths']:.3f})")
    conv = [r for r in rows if r["log_deaths"] < 0.05]
    d = np.array([r["ours"] - r["log500"] for r in conv])
    z = np.array([(r["ours"] - r["log500"]) / max(r["se"], 1.0) for r in conv])
    allr = np.corrcoef([r["ours"] for r in rows], [r["log500"] for r in rows])[0, 1]
    print(f"\n{len(rows)} models; correlation(ours, log last 500) over all models = {allr:.3f}")
    print(f"{len(conv)} converged runs (logged death rate < 5 %): ours - log(last 500): median {np.median(d):+.1f}, "
          f"mean {d.mean():+.1f}, |.| 90th percentile {np.percentile(np.abs(d), 90):.1f}; median |z| = {np.median(np.abs(z)):.2f}")
