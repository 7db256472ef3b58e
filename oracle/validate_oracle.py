"""Differential check: C oracle (oracle/wf_oracle.c) vs the UNMODIFIED Python reference.

TEST INFRASTRUCTURE; needs /root/reference (build container only).  Run as
    python -m oracle.validate_oracle [--steps N]
    python -m oracle.validate_oracle --random 200 --seed 1 [--steps N]   # randomised scenarios instead of the list
Every step compares type / burning / fm_inf / fuel / agent_pos planes, agent
xy + alive, fire_at_border, obs, reward (float64 ==) and done bit-exactly, and
temp on grass cells to 1e-12.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np

from . import wf_oracle as wo
from .policies import ring_actions
from .ref_harness import RefEnv

SCENARIOS = [
    dict(name="c1_10", width=10, height=10, seed=0),
    dict(name="c2_14", width=14, height=14, seed=1),
    dict(name="wind_random_14", width=14, height=14, seed=2, wind="random"),
    dict(name="wind_e_14", width=14, height=14, seed=3, wind=[0.85, (1, 0)]),
    dict(name="wind_diag_12", width=12, height=12, seed=4, wind=[0.85, (1, 1)]),
    dict(name="wind_sw_16", width=16, height=16, seed=12, wind=[0.85, (-1, 1)], extra_ignitions=3),
    dict(name="rivers_14", width=14, height=14, seed=5, make_rivers=True),
    dict(name="rivers_wind_20", width=20, height=20, seed=6, make_rivers=True, wind="random"),
    dict(name="digtoggle_10", width=10, height=10, seed=7, allow_dig_toggle=True, n_actions=6),
    dict(name="aspeed2_10", width=10, height=10, seed=8, a_speed=2),
    dict(name="aspeed3_toggle_12", width=12, height=12, seed=9, a_speed=3, allow_dig_toggle=True, n_actions=5),
    dict(name="ignite_32", width=32, height=32, seed=10, wind=[0.85, (1, 0)], extra_ignitions=6),
    dict(name="ignite_24_rand", width=24, height=24, seed=11, wind="random", extra_ignitions=4, make_rivers=True),
    dict(name="noop_actions_10", width=10, height=10, seed=13, n_actions=5),
    # scripted containment: ring walk, then random actions (latch, post-containment death / burn-out)
    dict(name="ring2_10", width=10, height=10, seed=20, policy="ring2"),
    dict(name="ring2_14", width=14, height=14, seed=21, policy="ring2"),
    dict(name="ring3_14_wind", width=14, height=14, seed=22, policy="ring3", wind="random"),
    dict(name="ring3_20_rivers", width=20, height=20, seed=23, policy="ring3", make_rivers=True),
    dict(name="ring2_12_toggle", width=12, height=12, seed=24, policy="ring2", allow_dig_toggle=True, n_actions=6),
    dict(name="ring4_16_aspeed2", width=16, height=16, seed=25, policy="ring4", a_speed=2),
    # the reference's heuristic walk policy (DQN.choose_randomwalk_action), its own code vs the C restatement
    dict(name="walk_10", width=10, height=10, seed=40, policy="walk"),
    dict(name="walk_14", width=14, height=14, seed=41, policy="walk"),
    dict(name="walk_20_rivers_wind", width=20, height=20, seed=42, policy="walk", make_rivers=True, wind="random"),
]


def random_scenario(rng, i, wide=False):
    """A random configuration of everything the step reads.  wide: W > H maps, where the literal [HEIGHT-1, y] border
    points of environment.py:222 are an interior column (the oracle keeps the reference's destructive deque there)."""
    size = int(rng.choice([10, 11, 12, 13, 14, 16, 18, 20, 24, 28]))
    width = size + (int(rng.integers(1, 7)) if wide else 0)
    sc = dict(name=f"rand{i}_{width}x{size}", width=width, height=size, seed=int(rng.integers(1, 1 << 30)))
    w = rng.integers(0, 4)
    if w == 1:
        sc["wind"] = "random"
    elif w == 2:
        sc["wind"] = [float(rng.choice([0.7, 0.85, 1.0])), (int(rng.integers(-1, 2)), int(rng.integers(-1, 2)))]
    if rng.random() < 0.3:
        sc["make_rivers"] = True
    if rng.random() < 0.3:
        sc["allow_dig_toggle"] = True
        sc["n_actions"] = int(rng.choice([5, 6]))
    elif rng.random() < 0.2:
        sc["n_actions"] = 5  # action 4 is a no-op without the toggle
    if rng.random() < 0.3:
        sc["a_speed"] = int(rng.choice([2, 3]))
    if rng.random() < 0.4:
        sc["extra_ignitions"] = int(rng.integers(1, 7))
    pol = rng.random()
    if pol < 0.25:
        sc["policy"] = f"ring{int(rng.integers(2, min(5, size // 2 - 1)))}"
    elif pol < 0.5 and "n_actions" not in sc:
        sc["policy"] = "walk"
    return sc


def compare(tag, ref: RefEnv, orc: wo.OracleEnv, obs_r, obs_o, rew=None, done=None):
    pr, po = ref.planes(), orc.planes()
    for k in ("type", "burning", "fm_inf", "fuel", "apos"):
        if not np.array_equal(pr[k], po[k]):
            raise AssertionError(f"{tag}: plane {k} differs\nref=\n{pr[k].T}\noracle=\n{po[k].T}")
    for k in ("alive", "ax", "ay", "fire_at_border", "running", "wind_x", "wind_y"):
        if pr[k] != po[k]:
            raise AssertionError(f"{tag}: scalar {k}: ref {pr[k]} oracle {po[k]}")
    assert pr["wind_speed"] == po["wind_speed"], tag
    g = pr["type"] == 0
    err = np.abs(pr["temp"] - po["temp"])[g].max() if g.any() else 0.0
    assert err <= 1e-12, f"{tag}: temp err {err}"
    assert np.array_equal(obs_r.astype(np.uint8), obs_o), f"{tag}: obs"
    assert set(np.unique(obs_r)) <= {0.0, 1.0}
    if rew is not None:
        assert float(rew[0]) == float(rew[1]), f"{tag}: reward ref {rew[0]} oracle {rew[1]}"
        assert bool(done[0]) == bool(done[1]), f"{tag}: done"
    return err


def run(sc, steps):
    cfg = {k: v for k, v in sc.items() if k not in ("name", "policy")}
    policy = sc.get("policy", "random")
    ref, orc = RefEnv(cfg), wo.OracleEnv(cfg)
    o_r, o_o = ref.reset(), orc.reset()
    compare(sc["name"] + " reset0", ref, orc, o_r, o_o)
    stats = dict(steps=0, episodes=0, deaths=0, contained=0, burnouts=0, maxerr=0.0)

    def script():
        if not policy.startswith("ring"):
            return []
        p = ref.planes()
        return ring_actions(p["ax"], p["ay"], ref.W // 2, ref.H // 2, int(policy[4:]))

    plan = script()
    for s in range(steps):
        a = ref.random_action()
        assert a == orc.random_action()
        if ref.t < len(plan):
            a = plan[ref.t]
        if policy == "walk":
            a = ref.walk_action()
            assert a == orc.walk_action(), f"{sc['name']} ep{ref.episode} t{ref.t}: walk action"
        o_r, r_r, d_r, _ = ref.step(a)
        o_o, r_o, d_o, _ = orc.step(a)
        tag = f"{sc['name']} ep{ref.episode} t{ref.t} a{a}"
        stats["maxerr"] = max(stats["maxerr"], compare(tag, ref, orc, o_r, o_o, (r_r, r_o), (d_r, d_o)))
        stats["steps"] += 1
        if r_r == cfg.get("contained_bonus", 1000):
            stats["contained"] += 1
        if d_r:
            stats["episodes"] += 1
            if not ref.sim.W.agents:
                stats["deaths"] += 1
            else:
                stats["burnouts"] += 1
            o_r, o_o = ref.reset(), orc.reset()
            compare(sc["name"] + f" reset ep{ref.episode}", ref, orc, o_r, o_o)
            plan = script()
    return stats


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--only", default=None)
    ap.add_argument("--random", type=int, default=0, help="number of randomised scenarios to run instead of the list")
    ap.add_argument("--wide", action="store_true", help="with --random: W > H maps")
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args(argv)
    total = 0
    scenarios = SCENARIOS
    if args.random:
        rng = np.random.default_rng(args.seed)
        scenarios = [random_scenario(rng, i, args.wide) for i in range(args.random)]
    for sc in scenarios:
        if args.only and args.only != sc["name"]:
            continue
        st = run(sc, args.steps)
        total += st["steps"]
        print(sc["name"], st, flush=True)
    print("OK: oracle == reference on", total, "steps")


if __name__ == "__main__":
    sys.exit(main())
