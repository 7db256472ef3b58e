"""Write tests/golden/*.npz by running the UNMODIFIED Python reference.

TEST INFRASTRUCTURE; needs /root/reference (build container only):
    python -m oracle.gen_golden
Each fixture is a sequence of frames.  Frame 0 is the state after the first
``reset()``; every ``step(a)`` adds a frame; when ``done`` is returned the next
frame is the ``reset()`` that follows (kind 0).  Randomness: oracle/philox.py.

Fields (F frames, grid W x H, index [f, x, y]):
  kind[F]      0 = reset frame, 1 = step frame
  action[F]    action passed to step (-1 on reset frames)
  reward[F]    float64 return of World.get_reward (0 on reset frames)
  done[F]      not World.RUNNING
  type, burning, fm_inf, fuel, apos   uint8 planes   (environment.py layers / burning_cells)
  temp         float64 plane -- only meaningful where type == 0
  alive, ax, ay, fire_at_border, wind_x, wind_y, wind_speed   per-frame scalars
  cfg          JSON of the scenario (METADATA keys + seed / extra_ignitions / policy)
"""
from __future__ import annotations

import json
import os

import numpy as np

from .policies import ring_actions
from .ref_harness import RefEnv

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SCENARIOS = [
    # BASELINE.json configs[0]: single 10x10 env, Logs/10-sized constants, seed 0, 200 steps
    dict(name="c1_10x10_seed0", width=10, height=10, seed=0, steps=200),
    dict(name="c2_14x14_seed1", width=14, height=14, seed=1, steps=400),
    dict(name="ring2_10x10", width=10, height=10, seed=20, policy="ring2", steps=400),
    dict(name="ring2_14x14", width=14, height=14, seed=21, policy="ring2", steps=500),
    dict(name="ring3_14x14_windrandom", width=14, height=14, seed=22, policy="ring3", wind="random", steps=400),
    dict(name="wind_e085_14x14", width=14, height=14, seed=3, wind=[0.85, (1, 0)], steps=300),
    dict(name="wind_diag_12x12", width=12, height=12, seed=4, wind=[0.85, (1, 1)], steps=200),
    dict(name="rivers_14x14", width=14, height=14, seed=5, make_rivers=True, steps=400),
    dict(name="rivers_ring3_20x20_windrandom", width=20, height=20, seed=23, policy="ring3", make_rivers=True,
         wind="random", steps=400),
    dict(name="digtoggle_10x10", width=10, height=10, seed=7, allow_dig_toggle=True, n_actions=6, steps=300),
    dict(name="aspeed2_ring4_16x16", width=16, height=16, seed=25, policy="ring4", a_speed=2, steps=400),
    dict(name="ignite6_wind_32x32", width=32, height=32, seed=10, wind=[0.85, (1, 0)], extra_ignitions=6, steps=300),
    dict(name="ignite4_rivers_24x24_windrandom", width=24, height=24, seed=11, wind="random", extra_ignitions=4,
         make_rivers=True, steps=300),
    # larger than one 32-lane bitboard row: exercises the tiled large-grid kernel
    dict(name="ignite8_wind_48x48", width=48, height=48, seed=30, wind=[0.85, (1, 0)], extra_ignitions=8, steps=150),
    dict(name="ring5_40x40", width=40, height=40, seed=31, policy="ring5", steps=250),
    dict(name="ignite12_rivers_64x64_windrandom", width=64, height=64, seed=32, wind="random", extra_ignitions=12,
         make_rivers=True, steps=120),
]


def record(sc):
    cfg = {k: v for k, v in sc.items() if k not in ("name", "policy", "steps")}
    policy = sc.get("policy", "random")
    ref = RefEnv(cfg)
    frames = []

    def snap(kind, action, reward, done):
        p = ref.planes()
        frames.append(dict(kind=kind, action=action, reward=float(reward), done=int(done), **p))

    def script():
        if not policy.startswith("ring"):
            return []
        p = ref.planes()
        return ring_actions(p["ax"], p["ay"], ref.W // 2, ref.H // 2, int(policy[4:]))

    ref.reset()
    snap(0, -1, 0.0, 0)
    plan = script()
    for _ in range(sc["steps"]):
        a = ref.random_action()
        if ref.t < len(plan):
            a = plan[ref.t]
        _, r, d, _ = ref.step(a)
        snap(1, a, r, d)
        if d:
            ref.reset()
            snap(0, -1, 0.0, 0)
            plan = script()
    out = dict(cfg=np.array(json.dumps({k: v for k, v in sc.items() if k != "name"})))
    for k in ("kind", "action", "done", "alive", "ax", "ay", "fire_at_border", "running", "wind_x", "wind_y"):
        out[k] = np.array([f[k] for f in frames], np.int32)
    out["reward"] = np.array([f["reward"] for f in frames], np.float64)
    out["wind_speed"] = np.array([f["wind_speed"] for f in frames], np.float64)
    for k in ("type", "burning", "fm_inf", "fuel", "apos"):
        out[k] = np.stack([f[k] for f in frames]).astype(np.uint8)
    temp = np.stack([f["temp"] for f in frames])
    temp[out["type"] != 0] = 0.0  # order-dependent / never read again in the reference: not part of the contract
    out["temp"] = temp
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for sc in SCENARIOS:
        out = record(sc)
        path = os.path.join(OUT, sc["name"] + ".npz")
        np.savez_compressed(path, **out)
        n_done = int(out["done"].sum())
        n_cont = int((out["reward"] == 1000).sum())
        print(f"{sc['name']}: {len(out['kind'])} frames, {n_done} episodes ended, {n_cont} containment rewards, "
              f"{os.path.getsize(path) / 1024:.1f} KiB", flush=True)


if __name__ == "__main__":
    main()
