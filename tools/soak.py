"""Randomised differential soak test: CUDA path vs the C oracle on random configurations.

    python tools/soak.py [seconds] [seed] [rollout|api|mlp]

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


def api_round(rng):
    """The per-call API instead of fused rollouts: step / step_host with explicit actions, masked resets as soon as
    envs finish, World.set_fire_to injections, and a checkpoint round trip (get_state -> set_state into a second
    handle that then has to stay in lock-step)."""
    fam, cfg = draw_config(rng)
    if fam == "tile":
        T, CS = rng.choice([(0, 0), (128, 1), (128, 2), (256, 2), (128, 4), (256, 8), (128, 16)])
        os.environ["WF_TILE_T"], os.environ["WF_TILE_CS"] = str(int(T)), str(int(CS))
    os.environ["WF_HOST_THREADS"] = str(int(rng.integers(1, 6)))
    N = int(rng.integers(2, 20)) if fam == "warp" else int(rng.integers(2, 6))
    K = int(rng.integers(40, 160)) if fam == "warp" else int(rng.integers(30, 80))
    print(f"  next(api): {fam} {cfg} N={N} K={K} T={os.environ.get('WF_TILE_T')} CS={os.environ.get('WF_TILE_CS')}", flush=True)
    W, H = cfg["width"], cfg["height"]
    gpu = BatchedForestFire(N, **cfg)
    twin = None
    orc = [wo.OracleEnv(cfg, env_id=i) for i in range(N)]
    # step_host through the resident step server on some rounds (warp family; the tile family refuses), with either
    # transport and a short idle limit, so that it parks itself / is parked by the device-side calls mixed in below
    os.environ["WF_SESSION_SECTORS"] = str(int(rng.integers(0, 2)))
    os.environ["WF_SESSION_IDLE_US"] = str(int(rng.choice([50, 300, 2000])))
    persistent = bool(rng.random() < 0.5)  # wf_host_session mode 2: change-list records, the array is patched in place
    session = bool(rng.random() < 0.7) and gpu.host_session(True, persistent_obs=persistent)
    print(f"            session={session} persistent_obs={persistent} sectors={os.environ['WF_SESSION_SECTORS']} "
          f"idle_us={os.environ['WF_SESSION_IDLE_US']}", flush=True)
    obs = to_np(gpu.reset())
    for i, e in enumerate(orc):
        assert np.array_equal(obs[i], e.reset()), "reset obs"
    a_speed = cfg.get("a_speed", 1)
    a_iter = a_speed
    n_act = cfg.get("n_actions", 4)
    for k in range(K):
        op = rng.random()
        if op < 0.08:  # World.set_fire_to on a few envs
            cells = np.full((N, 2), -1, np.int32)
            for i in rng.choice(N, size=max(1, N // 3), replace=False):
                cells[i] = (int(rng.integers(0, W)), int(rng.integers(0, H)))
                orc[i].set_fire_to(int(cells[i, 0]), int(cells[i, 1]))
            gpu.set_fire_to(torch.from_numpy(cells))
            if twin is not None:
                twin.set_fire_to(torch.from_numpy(cells))
        if op > 0.95 and twin is None:  # checkpoint into a second handle
            st = gpu.get_state()
            twin = BatchedForestFire(N, **cfg)
            twin.set_state(**st)
        acts = rng.integers(0, n_act + 1, size=N).astype(np.int32)
        live = [bool(e.planes()["running"]) for e in orc]
        for e in orc:
            e.set_a_speed_iter(a_iter)
        a_iter = a_speed if a_iter == 1 else a_iter - 1
        if rng.random() < 0.5:
            o_g, r_g, d_g, _ = gpu.step_host(acts)
            o_g, r_g, d_g = np.array(o_g), np.array(r_g), np.array(d_g)
        else:
            o_g, r_g, d_g, _ = gpu.step(torch.from_numpy(acts).cuda())
            o_g, r_g, d_g = to_np(o_g), to_np(r_g), to_np(d_g)
        for i, e in enumerate(orc):
            if not live[i]:
                assert r_g[i] == 0.0 and d_g[i], (i, k, "frozen env")
                continue
            o, r, d, _ = e.step(int(acts[i]))
            assert r_g[i] == r and bool(d_g[i]) == d, (i, k, r_g[i], r, d_g[i], d)
            assert np.array_equal(o_g[i], o), (i, k, "obs")
        if twin is not None:  # (the checkpoint carries a_speed_iter)
            o_t, r_t, d_t, _ = twin.step(torch.from_numpy(acts).cuda())
            assert np.array_equal(to_np(o_t), o_g) and np.array_equal(to_np(r_t), r_g), (k, "twin diverged")
        m = np.array([not e.planes()["running"] for e in orc], np.uint8)
        if m.any() and rng.random() < 0.5:
            o_r = to_np(gpu.reset(mask=torch.from_numpy(m).cuda()))
            if twin is not None:
                twin.reset(mask=torch.from_numpy(m).cuda())
            for i in np.nonzero(m)[0]:
                assert np.array_equal(o_r[i], orc[i].reset()), (i, k, "masked reset obs")
    compare_states("end", gpu, orc)
    gpu.close()
    if twin is not None:
        twin.close()
    return N * K


def mlp_round(rng):
    """WF_POLICY_MLP with random networks: hidden size, action count, grid size (both lane layouts of the warp family),
    epsilon; every chosen action against a float64 evaluation, every transition against the oracle."""
    from oracle import philox
    size = int(rng.integers(10, 33))
    n_act = int(rng.choice([4, 4, 5, 6]))
    cfg = dict(width=size, height=size, seed=int(rng.integers(1, 1 << 30)), n_actions=n_act)
    if n_act > 4 and rng.random() < 0.7:
        cfg["allow_dig_toggle"] = True
    if rng.random() < 0.4:
        cfg["wind"] = "random"
    if rng.random() < 0.3:
        cfg["make_rivers"] = True
    hid = int(rng.integers(1, 65))
    eps = float(rng.choice([0.0, 0.05, 0.3, 1.0]))
    N, K = int(rng.integers(2, 30)), int(rng.integers(40, 200))
    print(f"  next(mlp): {cfg} hid={hid} eps={eps} N={N} K={K}", flush=True)
    scale = float(rng.choice([0.1, 1.0, 10.0]))
    w1 = (rng.standard_normal((size * size * 3, hid)) * scale / np.sqrt(size)).astype(np.float32)
    b1 = (rng.standard_normal(hid) * scale).astype(np.float32)
    w2 = (rng.standard_normal((hid, n_act)) * scale).astype(np.float32)
    b2 = (rng.standard_normal(n_act) * scale).astype(np.float32)
    gpu = BatchedForestFire(N, auto_reset=True, **cfg)
    gpu.set_policy_mlp(w1, b1, w2, b2, eps=eps)
    orc = [wo.OracleEnv(cfg, env_id=i) for i in range(N)]
    obs0 = to_np(gpu.reset())
    for e in orc:
        e.reset()
    obs, rew, done, acts = gpu.rollout(K, policy="mlp", return_actions=True)
    obs, rew, done, acts = to_np(obs), to_np(rew), to_np(done), to_np(acts)
    eps_u32 = 0xFFFFFFFF if eps >= 1.0 else int(eps * 4294967296.0)
    w1d, b1d, w2d, b2d = (a.astype(np.float64) for a in (w1, b1, w2, b2))
    ties = 0
    for i, e in enumerate(orc):
        episode, t = 0, 0
        for k in range(K):
            seen = obs0[i] if k == 0 else obs[k - 1, i]
            u0, u1 = philox.explore_draw(cfg["seed"], i, episode, t)
            a = int(acts[k, i])
            if u0 < eps_u32:
                assert a == u1 % n_act, (i, k, "explore")
            else:
                h = seen.reshape(-1).astype(np.float64) @ w1d + b1d
                q = (1.0 / (1.0 + np.exp(-np.clip(h, -80, 80)))) @ w2d + b2d
                assert q[a] >= q.max() - 1e-3 * max(1.0, np.abs(q).max()), (i, k, a, q)
                ties += int(a != int(np.argmax(q)))
            o, r, d, _ = e.step(a)
            assert rew[k, i] == r and bool(done[k, i]) == d, (i, k)
            t += 1
            if d:
                o = e.reset()
                episode, t = episode + 1, 0
            assert np.array_equal(obs[k, i], o), (i, k, "obs")
    assert ties <= max(2, N * K // 200), ("too many near-ties", ties)
    compare_states("end", gpu, orc)
    gpu.close()
    return N * K


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    mode = sys.argv[3] if len(sys.argv) > 3 else "rollout"
    t0, rounds, steps = time.time(), 0, 0
    while mode in ("api", "mlp") and time.time() - t0 < budget:
        steps += api_round(rng) if mode == "api" else mlp_round(rng)
        rounds += 1
        print(f"round {rounds}: ok", flush=True)
    while mode not in ("api", "mlp") and time.time() - t0 < budget:
        fam, cfg, N, K, policy, n_done = one_round(rng)
        rounds += 1
        steps += N * K
        print(f"round {rounds}: episodes={n_done} ok", flush=True)
    print(f"SOAK OK: {rounds} rounds, {steps} env-steps compared in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
