"""The host-buffer entry points (wf_step_host / wf_reset_host, include/wildfire.h) the way a maintainer of the
reference would call them: plain pageable NumPy arrays through the ctypes binding of INTEGRATION.md section 3
(kept verbatim in tools/reference_binding/forest_fire_b200.py), inside a stand-in ``Simulation`` package that
holds only the two modules the binding imports (constants.METADATA, utility.grass).  Plus the ordering and
bookkeeping contracts of the handle: *_host calls run behind earlier *_dev calls, the caller's current device is
left alone, wf_stats counts fire ticks, a checkpoint carries the tick phase (a_speed > 1)."""
import importlib
import os
import re
import shutil
import sys

import numpy as np
import pytest
import torch

from oracle import wf_oracle as wo
from tests.gpu_util import compare_states, make_pair, to_np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINDING = os.path.join(ROOT, "tools", "reference_binding", "forest_fire_b200.py")


def test_integration_md_snippet_is_the_tested_file():
    """INTEGRATION.md section 3 shows exactly the file the GPU test below executes (CPU-only check)."""
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    sec = md[md.index("## 3. The reference-side binding"):md.index("## 4.")]
    code = re.search(r"```python\n(.*?)```", sec, re.S).group(1)
    assert code == open(BINDING).read()


def _stand_in_package(tmp_path, metadata, grass):
    pkg = tmp_path / "Simulation"
    pkg.mkdir()
    (pkg / "__init__.py").write_text("")
    (pkg / "constants.py").write_text(f"METADATA = {metadata!r}\n")
    (pkg / "utility.py").write_text(f"grass = {grass!r}\n")
    shutil.copy(BINDING, pkg / "forest_fire_b200.py")
    from wildfire_control_python_b200 import _lib
    os.symlink(_lib.LIB_PATH, pkg / "libwildfire_b200.so")
    for name in [m for m in sys.modules if m == "Simulation" or m.startswith("Simulation.")]:
        del sys.modules[name]
    sys.path.insert(0, str(tmp_path))
    try:
        return importlib.import_module("Simulation.forest_fire_b200")
    finally:
        sys.path.remove(str(tmp_path))


@pytest.mark.gpu
@pytest.mark.parametrize("size,n_envs,wind,steps", [(14, 33, [0.54, (0, 0)], 80), (10, 5, [0.85, (1, 0)], 60), (40, 6, [0.54, (0, 0)], 50)],
                         ids=["14x14_n33", "10x10_wind", "tile_40x40"])
def test_integration_snippet_pageable_numpy_buffers_match_oracle(tmp_path, size, n_envs, wind, steps):
    """The INTEGRATION.md binding, verbatim: pageable NumPy buffers land on wf_step_host's staged branch."""
    from wildfire_control_python_b200.constants import make_metadata
    meta = make_metadata(width=size, height=size, wind=wind)
    metadata = {k: meta[k] for k in ("width", "height", "n_actions", "a_speed", "allow_dig_toggle", "make_rivers", "wind",
                                     "death_penalty", "contained_bonus", "default_reward")}
    metadata["wind"] = [wind[0], tuple(wind[1])]
    grass = {"heat": meta["heat"], "fuel": meta["fuel"], "threshold": meta["threshold"], "radius": 1}
    mod = _stand_in_package(tmp_path, metadata, grass)
    sim = mod.ForestFire(n_envs=n_envs, seed=77)
    cfg = dict(width=size, height=size, wind=wind, seed=77)
    orc = [wo.OracleEnv(cfg, env_id=i) for i in range(n_envs)]
    obs = sim.reset()
    assert obs.dtype == np.float64 and obs.shape == (n_envs, size, size, 3)
    for i, e in enumerate(orc):
        assert np.array_equal(obs[i], e.reset()), i
    assert not torch.as_tensor(sim.obs).is_pinned()  # really the pageable branch
    for s in range(steps):
        acts = [e.random_action() for e in orc]
        live = [bool(e.planes()["running"]) for e in orc]
        obs, rew, done, info = sim.step(acts)
        for i, e in enumerate(orc):
            if not live[i]:
                assert rew[i] == 0.0 and done[i]
                continue
            o, r, d, _ = e.step(acts[i])
            assert rew[i] == r and bool(done[i]) == d and np.array_equal(obs[i], o), (s, i)
    # masked reset of the finished envs through the host entry point, start cells drawn from the stream
    fin = np.array([0 if e.planes()["running"] else 1 for e in orc], np.uint8)
    if fin.any():
        rc = mod.lib.wf_reset_host(sim.h, fin.ctypes.data, None, sim.obs, 0)
        assert rc == 0
        for i, e in enumerate(orc):
            if fin[i]:
                assert np.array_equal(sim.obs[i], e.reset()), i
            else:
                assert np.array_equal(sim.obs[i], e.obs()), i


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,n_envs", [(dict(width=256, height=256, seed=31, extra_ignitions=16), 96),
                                        (dict(width=14, height=14, seed=32), 4096)], ids=["tile_256_n96", "warp_14_n4096"])
def test_step_host_is_ordered_behind_device_calls(cfg, n_envs):
    """reset() on the caller's stream, then step_host() at once (bench.py and every e2e loop do this): the step on
    the handle's private stream must see the finished reset.  A long kernel is queued in front of the reset so that
    an unordered step would certainly overtake it."""
    from wildfire_control_python_b200.batched import BatchedForestFire
    twin = BatchedForestFire(n_envs, **cfg)
    twin.reset()
    acts = np.random.default_rng(5).integers(0, 4, size=(4, n_envs), dtype=np.int32)
    ref = []
    for k in range(4):
        o, r, d, _ = twin.step(torch.from_numpy(acts[k]).cuda())
        ref.append((to_np(o).copy(), to_np(r).copy(), to_np(d).copy()))
    big = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
    for rep in range(3):
        gpu = BatchedForestFire(n_envs, **cfg)  # a fresh handle: its reset() opens episode 0, like the twin's
        for _ in range(6):
            big.normal_()  # ~ms of queued work on the current stream
        gpu.reset()
        for k in range(4):
            o, r, d, _ = gpu.step_host(acts[k])
            assert np.array_equal(o, ref[k][0]) and np.array_equal(r, ref[k][1]) and np.array_equal(d, ref[k][2]), (rep, k)
    # and the other direction: device-side calls after a host step see its result
    st_a, st_b = gpu.get_state(), twin.get_state()
    for key in ("type", "burning", "fuel", "scalars"):
        assert torch.equal(st_a[key][:, ...], st_b[key][:, ...]), key


@pytest.mark.gpu
@pytest.mark.parametrize("size", [14, 48], ids=["warp", "tile"])
def test_stats_count_fire_ticks(size):
    """wf_stats[5] (include/wildfire.h): env fire ticks = env-steps on which ForestFire.update ran (every a_speed steps)."""
    from wildfire_control_python_b200.batched import BatchedForestFire
    N, K = 7, 30
    for a_speed in (1, 3):
        env = BatchedForestFire(N, width=size, height=size, seed=3, a_speed=a_speed)
        env.reset(starts=torch.tensor([[size // 2 + 3, size // 2]] * N, dtype=torch.int32))
        env.rollout(K, actions=torch.full((K, N), 9, dtype=torch.int32, device="cuda"), obs=False)  # no-ops: nobody dies
        st = env.stats()
        assert st["env_steps"] == N * K
        assert st["ticks"] == N * (K // a_speed), (a_speed, st)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [dict(width=14, height=14, seed=821, a_speed=3), dict(width=48, height=40, seed=822, a_speed=2)],
                         ids=["warp_a3", "tile_a2"])
def test_checkpoint_restores_the_tick_phase(cfg):
    """a_speed > 1: METADATA['a_speed_iter'] (Q8) lives in the handle; get_state()/set_state() carry it, so a handle
    restored MID-PHASE ticks the fire on the same steps as the original."""
    from wildfire_control_python_b200.batched import BatchedForestFire
    N, K2 = 9, 60
    for K1 in (cfg["a_speed"] + 1, 2 * cfg["a_speed"] + 1):  # both leave the counter away from its initial value
        a = BatchedForestFire(N, auto_reset=True, **cfg)
        a.reset()
        gen = torch.Generator("cuda").manual_seed(cfg["seed"])
        acts = torch.randint(0, 4, (K1 + K2, N), dtype=torch.int32, device="cuda", generator=gen)
        a.rollout(K1, actions=acts[:K1], obs=False)
        st = a.get_state()
        assert int(st["a_iter"][0]) == a.a_speed_iter and a.a_speed_iter != cfg["a_speed"]
        b = BatchedForestFire(N, auto_reset=True, **cfg)
        b.set_state(**st)
        assert b.a_speed_iter == int(st["a_iter"][0])
        oa, ra, da = a.rollout(K2, actions=acts[K1:])
        ob, rb, db = b.rollout(K2, actions=acts[K1:])
        assert torch.equal(ra, rb) and torch.equal(da, db) and torch.equal(oa, ob)
        sa, sb = a.get_state(), b.get_state()
        for k in ("type", "burning", "fm_inf", "fuel", "apos"):
            assert torch.equal(sa[k], sb[k]), k
        assert torch.equal(sa["a_iter"], sb["a_iter"])


@pytest.mark.gpu
def test_set_fire_to_twice_counts_one_burning_cell():
    """World.set_fire_to adds to a SET (environment.py:236): re-igniting a burning cell leaves len(burning_cells) alone."""
    cfg = dict(width=14, height=14, seed=9)
    gpu, orc = make_pair(3, cfg)
    gpu.reset()
    for e in orc:
        e.reset()
    cells = torch.tensor([[2, 3], [-1, 0], [7, 7]], dtype=torch.int32)  # env 2: the centre already burns
    for _ in range(2):
        gpu.set_fire_to(cells)
    orc[0].set_fire_to(2, 3)
    orc[0].set_fire_to(2, 3)
    orc[2].set_fire_to(7, 7)
    compare_states("set_fire_to twice", gpu, orc, obs=gpu.observe())


@pytest.mark.gpu
def test_calls_leave_the_current_device_alone():
    from wildfire_control_python_b200.batched import BatchedForestFire
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    env = BatchedForestFire(8, device="cuda:1", width=14, height=14)
    env.reset()
    env.step(torch.zeros(8, dtype=torch.int32, device="cuda:1"))
    env.step_host(np.zeros(8, np.int32))
    assert torch.cuda.current_device() == 0


# ---------------------------------------------------------------------------------------------
# wf_host_session: the step kernel as a resident server driven through mapped host memory
@pytest.mark.gpu
@pytest.mark.parametrize("cfg,n_envs", [(dict(width=14, height=14, seed=601), 601), (dict(width=10, height=10, seed=602), 7),
                                        (dict(width=20, height=20, seed=603), 530), (dict(width=17, height=13, seed=604, wind="random", make_rivers=True), 45),
                                        (dict(width=32, height=32, seed=605, a_speed=2, allow_dig_toggle=True, n_actions=5), 19),
                                        (dict(width=14, height=14, seed=606, fuel=40, extra_ignitions=2), 64)],
                         ids=["14_n601", "10_n7", "20_n530", "17x13_rivers_windrandom", "32_aspeed2_toggle", "14_fuel40"])
@pytest.mark.parametrize("transport", ["flag", "sectors", "persistent"])
def test_host_session_matches_oracle(monkeypatch, cfg, n_envs, transport):
    """Every action / reward / done / observation of a session against the oracle, with auto-reset inside the resident
    kernel, and with the session interrupted by other entry points (which park the kernel) and by idle periods (after
    which it parks itself)."""
    import time
    monkeypatch.setenv("WF_HOST_THREADS", "5")
    monkeypatch.setenv("WF_SESSION_IDLE_US", "300")
    # how the records reach the host: one completion flag behind a system-scope fence (default), or self-validating
    # 32-byte sectors (seven payload words + sequence number ^ hash) that need neither
    monkeypatch.setenv("WF_SESSION_SECTORS", "1" if transport == "sectors" else "0")
    gpu, orc = make_pair(n_envs, cfg, auto_reset=True)
    gpu.reset()
    for e in orc:
        e.reset()
    # "persistent": wf_host_session mode 2 -- step_host hands out the same array every call, only the elements that
    # changed travel (change-list records; records with more than 14 changes, e.g. reset envs, travel in full)
    assert gpu.host_session(True, persistent_obs=transport == "persistent") and gpu.host_session_state == 1
    for s in range(150):
        acts = np.array([e.random_action() for e in orc], np.int32)
        obs, rew, done, _ = gpu.step_host(acts)
        assert gpu.host_session_state == 2
        for i, e in enumerate(orc):
            o, r, d, _ = e.step(int(acts[i]))
            if d:
                o = e.reset()
            assert rew[i] == r and bool(done[i]) == d, (s, i, rew[i], r)
            assert np.array_equal(obs[i], o), (s, i)
        if s == 40:
            time.sleep(0.02)  # longer than the idle limit: the kernel parks itself, the next step starts it again
        if s == 70:
            compare_states("mid-session", gpu, orc)  # wf_get_state parks the kernel; state is whole in HBM
            assert gpu.host_session_state == 1
        if s == 100:  # a device-side step in between (parks, steps, and the session resumes behind it)
            acts = [e.random_action() for e in orc]
            o_g, r_g, d_g, _ = gpu.step(torch.tensor(acts, dtype=torch.int32, device="cuda"))
            o_g, r_g, d_g = to_np(o_g), to_np(r_g), to_np(d_g)
            for i, e in enumerate(orc):
                o, r, d, _ = e.step(acts[i])
                if d:
                    o = e.reset()
                assert r_g[i] == r and bool(d_g[i]) == d and np.array_equal(o_g[i], o), i
    assert gpu.host_threads == 5
    st = gpu.stats()
    assert st["env_steps"] == 151 * n_envs
    assert gpu.host_session(False) is False and gpu.host_session_state == 0
    compare_states("after the session", gpu, orc, obs=gpu.observe())


@pytest.mark.gpu
def test_host_session_pageable_buffers_and_plain_path_agree():
    """The session needs no page-locked caller buffers, and gives exactly what the launch-per-step path gives."""
    from wildfire_control_python_b200 import _lib
    from wildfire_control_python_b200.batched import BatchedForestFire
    N, cfg = 300, dict(width=14, height=14, seed=611, auto_reset=True)
    a, b = BatchedForestFire(N, **cfg), BatchedForestFire(N, **cfg)
    a.reset(); b.reset()
    L = _lib.lib()
    assert L.wf_host_session(a._h, 1) == 0
    obs = np.zeros((N, 14, 14, 3), np.uint8); rew = np.zeros(N); done = np.zeros(N, np.uint8)  # pageable
    rng = np.random.default_rng(1)
    for s in range(60):
        acts = rng.integers(0, 4, N, dtype=np.int32)
        assert L.wf_step_host(a._h, acts.ctypes.data, obs.ctypes.data, 0, rew.ctypes.data, done.ctypes.data) == 0
        o2, r2, d2, _ = b.step_host(acts)
        assert np.array_equal(obs, o2) and np.array_equal(rew, r2) and np.array_equal(done.astype(bool), d2), s


@pytest.mark.gpu
def test_host_session_persistent_obs_follows_the_callers_array():
    """Mode 2 patches the caller's array in place; a different array than the previous call's (here: two arrays in turn,
    then one array for a while, then a fresh one) gets every record in full, so each array handed in is complete."""
    from wildfire_control_python_b200 import _lib
    from wildfire_control_python_b200.batched import BatchedForestFire
    N, cfg = 333, dict(width=14, height=14, seed=612, auto_reset=True)
    a, b = BatchedForestFire(N, **cfg), BatchedForestFire(N, **cfg)
    a.reset(); b.reset()
    L = _lib.lib()
    assert L.wf_host_session(a._h, 2) == 0
    bufs = [np.full((N, 14, 14, 3), 7, np.uint8) for _ in range(3)]
    rew = np.zeros(N); done = np.zeros(N, np.uint8)
    rng = np.random.default_rng(2)
    for s in range(120):
        obs = bufs[s & 1] if s < 40 else bufs[0] if s < 100 else bufs[2]
        if s == 100:
            bufs[2][:] = 9  # never seen by the library: must come back complete
        acts = rng.integers(0, 4, N, dtype=np.int32)
        assert L.wf_step_host(a._h, acts.ctypes.data, obs.ctypes.data, 0, rew.ctypes.data, done.ctypes.data) == 0
        o2, r2, d2, _ = b.step_host(acts)
        assert np.array_equal(obs, o2) and np.array_equal(rew, r2) and np.array_equal(done.astype(bool), d2), s
    assert L.wf_host_session(a._h, 1) == 0  # back to full records, same array: nothing stale
    for s in range(10):
        acts = rng.integers(0, 4, N, dtype=np.int32)
        assert L.wf_step_host(a._h, acts.ctypes.data, bufs[2].ctypes.data, 0, rew.ctypes.data, done.ctypes.data) == 0
        o2, r2, d2, _ = b.step_host(acts)
        assert np.array_equal(bufs[2], o2) and np.array_equal(rew, r2), s
    assert L.wf_host_session(a._h, 3) == _lib.WF_ERR_INVALID


@pytest.mark.gpu
def test_host_session_is_refused_for_the_tile_family():
    from wildfire_control_python_b200.batched import BatchedForestFire
    env = BatchedForestFire(4, width=48, height=48)
    env.reset()
    assert env.host_session(True) is False and env.host_session_state == 0
    env.step_host(np.zeros(4, np.int32))  # the launch-per-step path still serves it


@pytest.mark.gpu
def test_host_session_falls_back_when_the_batch_cannot_be_resident():
    """The step server is a cooperative launch (every CTA resident).  A batch with more CTAs than the GPU has slots cannot
    be one: the first step_host turns the session off and the launch-per-step path serves the call -- same results."""
    from wildfire_control_python_b200.batched import BatchedForestFire
    N, cfg = 6000, dict(width=10, height=10, seed=621, auto_reset=True)  # 750 CTAs of 8 envs > 592 slots
    a, b = BatchedForestFire(N, **cfg), BatchedForestFire(N, **cfg)
    a.reset(); b.reset()
    assert a.host_session(True)
    rng = np.random.default_rng(2)
    for s in range(12):
        acts = rng.integers(0, 4, N, dtype=np.int32)
        oa, ra, da, _ = a.step_host(acts)
        ob, rb, db, _ = b.step_host(acts)
        assert np.array_equal(oa, ob) and np.array_equal(ra, rb) and np.array_equal(da, db), s
    assert a.host_session_state == 0
