"""Next-row N3: the reference's heuristic Baseline policy.

`DQN.choose_randomwalk_action` (DQN.py:353-389) drives `collect_memories(perform_baseline=True)`
(DQN.py:286-348).  The thesis reports its mean episode return over the last 2500 of 10 000 episodes:
1129 on 10x10 and 1152 on 14x14 (Report/results.tex:30,112; Plots/results_second.txt:2).
Reproducing those two numbers pins the WHOLE path -- fire timing, containment search, reward
arithmetic, start-cell sampling, policy -- to the reference's published results.
The C restatement of the policy is checked against the reference's own code in
oracle/validate_oracle.py (scenarios walk_*).
"""
import numpy as np
import pytest
import torch

from oracle import wf_oracle as wo

REPORTED = {10: 1129.0, 14: 1152.0}


def episode_returns(reward, done, start_before):
    """Total reward of every episode that STARTS before step `start_before` (so that the sample is
    not biased towards short episodes by the end of the rollout); each must finish inside the rollout."""
    reward, done = np.asarray(reward, np.float64), np.asarray(done, bool)
    out = []
    for i in range(reward.shape[1]):
        ends = np.nonzero(done[:, i])[0]
        start = 0
        for e in ends:
            if start < start_before:
                out.append(reward[start:e + 1, i].sum())
            start = e + 1
        assert start >= start_before, "an episode that started early did not finish: lengthen the rollout"
    return np.array(out)


@pytest.mark.parametrize("size", [10, 14])
def test_oracle_baseline_policy_matches_reported_mean_return(size):
    env = wo.OracleEnv(dict(width=size, height=size, seed=2019))
    rets = []
    for _ in range(2500):
        env.reset()
        total, done = 0.0, False
        while not done:
            _, r, done, _ = env.step(env.walk_action())
            total += r
        rets.append(total)
    rets = np.array(rets)
    se = rets.std() / np.sqrt(len(rets))
    assert se < 12
    assert abs(rets.mean() - REPORTED[size]) < 4 * se + 5, (rets.mean(), se)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [dict(width=14, height=14, seed=71), dict(width=10, height=10, seed=72, wind="random"),
                                 dict(width=40, height=36, seed=73)], ids=["14", "10_wind", "tile_40x36"])
def test_gpu_walk_policy_matches_oracle(cfg):
    from tests.gpu_util import compare_states, make_pair, to_np
    N, K = 41, 260
    gpu, orc = make_pair(N, cfg, auto_reset=True)
    gpu.reset()
    for e in orc:
        e.reset()
    obs, rew, done, acts = gpu.rollout(K, policy="walk", return_actions=True)
    obs, rew, done, acts = to_np(obs), to_np(rew), to_np(done), to_np(acts)
    n_contained = 0
    for i, e in enumerate(orc):
        for k in range(K):
            a = e.walk_action()
            assert acts[k, i] == a, f"env {i} step {k}: action {acts[k, i]} != {a}"
            o, r, d, _ = e.step(a)
            assert rew[k, i] == r and bool(done[k, i]) == d, (i, k)
            n_contained += int(r == 1000)
            if d:
                o = e.reset()
            assert np.array_equal(obs[k, i], o), (i, k)
    assert n_contained > N
    compare_states("after walk rollout", gpu, orc)


@pytest.mark.gpu
@pytest.mark.parametrize("size", [10, 14])
def test_gpu_baseline_reproduces_reported_mean_return(size):
    """~50 000 episodes on the GPU: the standard error is ~3, the thesis numbers are 1129 / 1152."""
    from wildfire_control_python_b200 import BatchedForestFire
    env = BatchedForestFire(4096, width=size, height=size, seed=5, auto_reset=True)
    env.reset()
    _, rew, done = env.rollout(1500, policy="walk", obs=False)
    rets = episode_returns(rew.cpu().numpy(), done.cpu().numpy(), start_before=1000)
    assert len(rets) > 30000
    se = rets.std() / np.sqrt(len(rets))
    assert abs(rets.mean() - REPORTED[size]) < 4 * se + 8, (rets.mean(), se, len(rets))
