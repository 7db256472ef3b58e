"""Sample (final ASCII map, total reward) pairs from the reference's own training logs.

TEST INFRASTRUCTURE; needs /root/reference (build container only):
    python -m oracle.gen_logs_kat
Logs/10-sized/* and Logs/14-sized/* were written by DQN.py:392-424; `maps` holds
[episode, ForestFire.render() string] and `total_rewards` the episode return.
For a non-death episode the reference's reward arithmetic (environment.py:342-390)
implies   total = 1000 + 1000 * (#'+') / N^2 - (T - 2)   with T the episode length:
T-2 steps at -1, one step at +1000 (containment, paid once) and the last step at
1000 * healthy / N^2 (burn-out).  tests/test_logs_kat.py replays each sampled
final dirt layout as a free burn and checks T and the burnt set.
"""
from __future__ import annotations

import glob
import json
import os

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "logs_kat.json")
REF = os.environ.get("WF_REFERENCE_ROOT", "/root/reference")


def main(per_size=300):
    out = {"source": "Logs/{10,14}-sized/*: maps[i] + total_rewards[i], non-death episodes, evenly sampled",
           "entries": []}
    for size in (10, 14):
        files = sorted(glob.glob(os.path.join(REF, "Logs", f"{size}-sized", "*")))
        pool = []
        for f in files:
            d = json.load(open(f))
            assert d["metadata"]["width"] == size
            for idx, m in d["maps"]:
                if not d["agent_deaths"][idx]:
                    pool.append(dict(size=size, file=os.path.basename(f), episode=idx, map=m,
                                     total_reward=d["total_rewards"][idx]))
        stride = max(1, len(pool) // per_size)
        picked = pool[::stride][:per_size]
        print(size, "pool", len(pool), "picked", len(picked))
        out["entries"] += picked
    json.dump(out, open(OUT, "w"))
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
