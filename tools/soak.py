"""Randomised differential soak test: CUDA path vs the C oracle on random configurations.

    python tools/soak.py [seconds] [seed]

Each round draws a configuration (grid size from both kernel families, wind fixed / random / directional,
rivers, dig toggle, a_speed, extra ignitions, fuel / threshold, tile cluster geometry), runs a fused
rollout with auto-reset under one of the device policies (ACTION stream, walk, explicit actions) and
replays every env on the oracle: actions, rewards, dones and observations of every step, then the
full state.  Prints one line per round; exits non-zero on the first mismatch."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wf_oracle as wo  # noqa: E402
from tests.gpu_util import compare_states, to_np  # noqa: E402
from wildfire_control_python_b200.batched import BatchedForestFire  # noqa: E402


def draw_config(rng):
    fam = rng.choice(["warp", "warp", "tile"])
    # W > H maps put the reference's literal border points [HEIGHT-1, y] on a column inside the map; once the fire
    # reaches that column the reference's answer depends on which border points its destructive search has already
    # thrown away, so there is no history-free rule to compare against: keep the fire origin
    # (x = W // 2) at least 30 cells from that column, or use square maps.
    if fam == "warp":
        W = H = int(rng.integers(10, 33))
    else:
        W = int(rng.choice([33, 40, 48, 64, 70, 96, 128, 130, 200]))
        H = int(rng.choice([h for h in (20, 33, 36, 40, 64, 70, 96, 100, 128) if h <= W and (h == W or abs(W // 2 - (h - 1)) >= 30)] or [W]))
    cfg = dict(width=W, height=H, seed=int(rng.integers(1, 1 << 30)))
    wind = rng.choice(["default", "random", "dir", "dir"])
    if wind == "random":
        cfg["wind"] = "random"
    elif wind == "dir":
        cfg["wind"] = [float(rng.choice([0.54, 0.7, 0.85])), (int(rng.integers(-1, 2)), int(rng.integers(-1, 2)))]
    if rng.random() < 0.35:
        cfg["make_rivers"] = True
    if rng.random() < 0.25:
        cfg["allow_dig_toggle"], cfg["n_actions"] = True, 6
    if rng.random() < 0.25:
        cfg["a_speed"] = int(rng.integers(2, 4))
    if rng.random() < 0.5:
        cfg["extra_ignitions"] = int(rng.integers(1, 9))
    if rng.random() < 0.2:
        cfg["fuel"] = int(rng.choice([8, 31, 40]))
        cfg["threshold"] = float(rng.choice([2.0, 3.0, 4.5]))
    return fam, cfg


def one_round(rng):
    fam, cfg = draw_config(rng)
    if fam == "tile":
        T, CS = rng.choice([(0, 0), (128, 1), (128, 2), (256, 2), (128, 4), (256, 8), (128, 16)][: 7])
        os.environ["WF_TILE_T"], os.environ["WF_TILE_CS"] = str(int(T)), str(int(CS))
    N = int(rng.integers(3, 40)) if fam == "warp" else int(rng.integers(2, 9))
    K = int(rng.integers(60, 260)) if fam == "warp" else int(rng.integers(40, 140))
    policy = rng.choice(["stream", "walk", "given"])
    print(f"  next: {fam} {cfg} N={N} K={K} policy={policy} T={os.environ.get('WF_TILE_T')} CS={os.environ.get('WF_TILE_CS')}", flush=True)
    gpu = BatchedForestFire(N, auto_reset=True, **cfg)
    orc = [wo.OracleEnv(cfg, env_id=i) for i in range(N)]
    obs0 = to_np(gpu.reset())
    for i, e in enumerate(orc):
        assert np.array_equal(obs0[i], e.reset()), "reset obs"
    given = None
    if policy == "given":
        given = torch.randint(0, cfg.get("n_actions", 4) + 1, (K, N), dtype=torch.int32, device="cuda")  # incl. a no-op id
        obs, rew, done = gpu.rollout(K, actions=given)
        acts = to_np(given)
    else:
        obs, rew, done, acts = gpu.rollout(K, policy=policy, return_actions=True)
        acts = to_np(acts)
    obs, rew, done = to_np(obs), to_np(rew), to_np(done)
    n_done = 0
    for i, e in enumerate(orc):
        for k in range(K):
            if policy == "stream":
                assert acts[k, i] == e.random_action(), (i, k, "stream action")
            elif policy == "walk":
                assert acts[k, i] == e.walk_action(), (i, k, "walk action")
            o, r, d, _ = e.step(int(acts[k, i]))
            assert rew[k, i] == r and bool(done[k, i]) == d, (i, k, rew[k, i], r, done[k, i], d)
            if d:
                o = e.reset()
                n_done += 1
            assert np.array_equal(obs[k, i], o), (i, k, "obs")
    compare_states("end", gpu, orc)
    gpu.close()
    return fam, cfg, N, K, policy, n_done


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    t0, rounds, steps = time.time(), 0, 0
    while time.time() - t0 < budget:
        fam, cfg, N, K, policy, n_done = one_round(rng)
        rounds += 1
        steps += N * K
        print(f"round {rounds}: episodes={n_done} ok", flush=True)
    print(f"SOAK OK: {rounds} rounds, {steps} env-steps compared in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
