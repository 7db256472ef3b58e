"""wf_get_state / wf_set_state / wf_set_fire_to (include/wildfire.h): checkpoint-and-restore of a running
batch and World.set_fire_to injection, for both kernel families and both hit-counter layouts.

The reference never checkpoints its environment; a batched env needs it (parity
injection, resuming rollouts).  The contract tested here: a handle restored from another handle's
exported state continues on EXACTLY the same trajectory."""
import numpy as np
import pytest
import torch

from tests.gpu_util import compare_states, make_pair, to_np

pytestmark = pytest.mark.gpu

CFGS = [
    dict(width=14, height=14, seed=801),                                              # warp family, bit-sliced hit totals
    dict(width=16, height=12, seed=802, wind="random", make_rivers=True),             # warp family, per-direction counters
    dict(width=64, height=64, seed=803, make_rivers=True, extra_ignitions=4),         # tile family, 2 words per row
    dict(width=128, height=128, seed=804, wind=[0.85, (1, 0)], extra_ignitions=8),    # tile family, 128-bit path
]


@pytest.mark.parametrize("cfg", CFGS, ids=lambda c: f"{c['width']}x{c['height']}_s{c['seed']}")
def test_restored_handle_continues_identically(cfg):
    from wildfire_control_python_b200.batched import BatchedForestFire
    N, K1, K2 = 9, 45, 70
    a = BatchedForestFire(N, auto_reset=True, **cfg)
    a.reset()
    gen = torch.Generator("cuda").manual_seed(cfg["seed"])
    acts = torch.randint(0, 4, (K1 + K2, N), dtype=torch.int32, device="cuda", generator=gen)
    a.rollout(K1, actions=acts[:K1], obs=False)
    st = a.get_state()
    b = BatchedForestFire(N, auto_reset=True, **cfg)  # never reset: everything comes from the checkpoint
    b.set_state(type=st["type"], burning=st["burning"], fm_inf=st["fm_inf"], fuel=st["fuel"], hits=st["hits"], scalars=st["scalars"])
    assert torch.equal(a.observe(), b.observe())
    oa, ra, da = a.rollout(K2, actions=acts[K1:])
    ob, rb, db = b.rollout(K2, actions=acts[K1:])
    assert torch.equal(ra, rb) and torch.equal(da, db) and torch.equal(oa, ob)
    if cfg["width"] <= 64:
        assert int(da.sum()) > 0  # resets (new Philox episodes) happened after the restore, too
    sa, sb = a.get_state(), b.get_state()
    for k in ("type", "burning", "fm_inf", "fuel", "apos"):
        assert torch.equal(sa[k], sb[k]), k
    assert torch.equal(sa["hits"].int().sum(-1), sb["hits"].int().sum(-1))
    assert torch.equal(sa["scalars"][:, :15], sb["scalars"][:, :15])


@pytest.mark.parametrize("cfg", [dict(width=14, height=14, seed=811), dict(width=48, height=40, seed=812, wind=[0.85, (0, 1)])],
                         ids=["warp", "tile"])
def test_set_fire_to_matches_oracle(cfg):
    """World.set_fire_to(cell) after reset() (environment.py:233-246) on a few envs, then stepping."""
    N = 6
    gpu, orc = make_pair(N, cfg)
    gpu.reset()
    for e in orc:
        e.reset()
    W, H = cfg["width"], cfg["height"]
    cells = np.full((N, 2), -1, np.int32)
    for i in range(0, N, 2):
        cells[i] = (2 + i, H - 1 - i)  # near / on the border
        orc[i].set_fire_to(int(cells[i, 0]), int(cells[i, 1]))
    gpu.set_fire_to(torch.from_numpy(cells))
    compare_states("after set_fire_to", gpu, orc, obs=gpu.observe())
    for s in range(60):
        acts = [e.random_action() for e in orc]
        live = [bool(e.planes()["running"]) for e in orc]
        obs, rew, done, _ = gpu.step(torch.tensor(acts, dtype=torch.int32, device="cuda"))
        rew, done = to_np(rew), to_np(done)
        for i, e in enumerate(orc):
            if live[i]:
                _, r, d, _ = e.step(acts[i])
                assert rew[i] == r and bool(done[i]) == d, (s, i)
    compare_states("60 steps after set_fire_to", gpu, orc, obs=gpu.observe())
