"""Next-row N1: the torch restatement of the reference's learners
(DQN.py, DQN_SARSA.py, DQN_DUEL.py, DQN_BOTH.py).

CPU tests drive the agents with a tiny stand-in environment that has the reference's surface and check
the learning rule against a literal per-sample restatement of the reference's ``replay`` loops
(DQN.py:156-185, DQN_SARSA.py:103-132); GPU tests run them on the CUDA environment.
"""
import json
import os

import numpy as np
import pytest
import torch

from wildfire_control_python_b200 import agents as A
from wildfire_control_python_b200.compat import get_name
from wildfire_control_python_b200.constants import make_metadata


class _FakeAgent:
    def __init__(self, sim):
        self.sim = sim

    x = property(lambda self: self.sim.ax)
    y = property(lambda self: self.sim.ay)

    def fire_in_direction(self, a):
        return False


class _FakeWorld:
    WIDTH, HEIGHT, DEPTH = 10, 10, 3
    wind_speed, wind_vector = 0.54, (0, 0)

    def __init__(self, sim):
        self.sim = sim

    @property
    def agents(self):
        return [] if self.sim.dead else [_FakeAgent(self.sim)]


class FakeSim:
    """10x10 stand-in with the reference's surface: the agent walks; an episode ends after `horizon` steps,
    with the containment bonus if it made at least 6 moves to the east or north, else with death."""

    def __init__(self, horizon=8, seed=0):
        self.METADATA = make_metadata(width=10, height=10)
        self.n_actions, self.DEBUG, self.get_name = 4, 1, get_name
        self.W = _FakeWorld(self)
        self.horizon, self.rng = horizon, np.random.default_rng(seed)
        self.reset()

    def _obs(self):
        o = np.zeros((10, 10, 3))
        o[self.ax, self.ay, 0] = 1
        o[5, 5, 1] = 1
        o[:, :, 2] = 1
        return o

    def reset(self):
        self.ax, self.ay, self.t, self.good, self.dead = int(self.rng.integers(3, 7)), int(self.rng.integers(3, 7)), 0, 0, False
        if (self.ax, self.ay) == (5, 5):  # the reference never starts the agent on the fire origin
            self.ax = 4
        return self._obs()

    def step(self, a):
        dx, dy = [(0, -1), (0, 1), (1, 0), (-1, 0)][int(a)]
        self.ax, self.ay = int(np.clip(self.ax + dx, 0, 9)), int(np.clip(self.ay + dy, 0, 9))
        self.t += 1
        self.good += int(a in (0, 2))
        done = self.t >= self.horizon
        reward = -1
        if done:
            if self.good >= 6:
                reward = self.METADATA["contained_bonus"]
            else:
                reward, self.dead = self.METADATA["death_penalty"], True
        return [self._obs(), reward, done, {}]

    def render(self, print_map=False):
        return "\n" + "\n".join("+" * 10 for _ in range(10)) + "\n"


def _fill(agent, n, sarsa):
    sim = agent.sim
    rng = np.random.default_rng(1)
    s = sim.reset()
    for _ in range(n):
        a = int(rng.integers(0, 4))
        sp, r, d, _ = sim.step(a)
        if sarsa:
            agent.remember(s, a, r, sp, int(rng.integers(0, 4)), d)
        else:
            agent.remember(s, a, r, sp, d)
        s = sim.reset() if d else sp


@pytest.mark.parametrize("cls", [A.DQN, A.DQN_SARSA, A.DQN_DUEL, A.DQN_BOTH])
def test_network_shapes_match_keras_models(cls):
    ag = cls(FakeSim(), verbose=False, device="cpu")
    n_in = 10 * 10 * 3
    want = n_in * 50 + 50 + 50 * 4 + 4  # Flatten -> Dense(50) -> Dense(4), DQN.py:211-219
    if cls in (A.DQN_DUEL, A.DQN_BOTH):
        want += n_in * 50 + 50 + 50 * 1 + 1  # value stream, DQN_DUEL.py:31-32
    assert sum(p.numel() for p in ag.model.parameters()) == want
    q = ag.model(torch.rand(5, 10, 10, 3))
    assert q.shape == (5, 4)
    for k, v in ag.get_weights().items():  # Keras orientation [in, out]
        if k == "dense_1/kernel:0":
            assert v.shape == (n_in, 50)
    assert all(torch.equal(a, b) for a, b in zip(ag.model.state_dict().values(), ag.target.state_dict().values()))


def test_dueling_head_is_v_plus_centred_advantage():
    ag = A.DQN_DUEL(FakeSim(), verbose=False, device="cpu")
    x = torch.rand(7, 10, 10, 3)
    m = ag.model
    adv = m.dense_2(torch.sigmoid(m.dense_1(x.flatten(1))))
    val = m.dense_4(torch.sigmoid(m.dense_3(x.flatten(1))))
    assert torch.allclose(m(x), val + adv - adv.mean(1, keepdim=True), atol=1e-6)


@pytest.mark.parametrize("cls,sarsa", [(A.DQN, False), (A.DQN_SARSA, True), (A.DQN_DUEL, False), (A.DQN_BOTH, True)])
def test_replay_targets_follow_the_reference_loop(cls, sarsa):
    """Per-sample restatement of DQN.py:164-180 / DQN_SARSA.py:110-125 vs the batched tensor version."""
    ag = cls(FakeSim(), verbose=False, device="cpu")
    with torch.no_grad():  # target != model, so that using the wrong network would show
        for p in ag.target.parameters():
            p.add_(0.05 * torch.randn_like(p))
    _fill(ag, 200, sarsa)
    batch = ag.memory.sample(32)
    got = ag.replay_targets(batch).numpy()
    s, a, r, sp, ap, d = [t.numpy() for t in batch]
    for i in range(32):
        with torch.no_grad():
            prediction = ag.target(torch.as_tensor(s[i:i + 1]).float())[0].numpy().copy()
            q_next = ag.target(torch.as_tensor(sp[i:i + 1]).float())[0].numpy()
        if d[i]:
            prediction[a[i]] = r[i]
        else:
            predQ = q_next[ap[i]] if sarsa else np.amax(q_next)
            prediction[a[i]] = r[i] + ag.gamma * predQ
        assert np.allclose(got[i], prediction, rtol=1e-5, atol=1e-4), i


def test_replay_step_is_clipped_adam_on_mse():
    ag = A.DQN(FakeSim(), verbose=False, device="cpu")
    _fill(ag, 100, False)
    before = [p.detach().clone() for p in ag.model.parameters()]
    loss = ag.replay()
    assert np.isfinite(loss) and loss > 0
    # rewards of +-1000 give huge gradients: with clipvalue=1 the first Adam step moves every weight by <= lr
    for b, p in zip(before, ag.model.parameters()):
        assert float((p.detach() - b).abs().max()) <= ag.alpha * 1.0001
    assert any(float((p.detach() - b).abs().max()) > 0 for b, p in zip(before, ag.model.parameters()))


def test_memory_behaves_like_a_bounded_deque():
    m = A.ReplayMemory((2, 2, 3), "cpu", maxlen=5)
    for k in range(8):
        m.append(np.full((2, 2, 3), k), k % 4, float(k), np.full((2, 2, 3), k + 1), k == 7)
    assert len(m) == 5
    assert sorted(float(x) for x in m._r[:5]) == [3.0, 4.0, 5.0, 6.0, 7.0]  # the three oldest were dropped
    s, a, r, sp, ap, d = m.sample(5)
    assert sorted(r.tolist()) == [3.0, 4.0, 5.0, 6.0, 7.0] and int(d.sum()) == 1
    assert torch.equal(sp, s + 1)
    grow = A.ReplayMemory((2, 2, 3), "cpu", maxlen=None)
    for k in range(5000):
        grow.append(np.zeros((2, 2, 3)), 0, float(k), np.zeros((2, 2, 3)), False)
    assert len(grow) == 5000 and float(grow._r[4999]) == 4999.0


def test_epsilon_schedule_and_greedy_choice():
    ag = A.DQN(FakeSim(), verbose=False, device="cpu")
    ag.decay_epsilon(100)
    assert ag.eps == pytest.approx(0.01 + 0.99 * np.exp(-0.005 * 100))  # DQN.py:199-202
    state = ag.sim.reset()
    with torch.no_grad():
        best = int(ag.model(torch.as_tensor(state).float()[None]).argmax())
    assert all(ag.choose_action(state, eps=0) == best for _ in range(5))
    assert {ag.choose_action(state, eps=1.0) for _ in range(200)} == {0, 1, 2, 3}


@pytest.mark.parametrize("cls", [A.DQN, A.DQN_SARSA])
def test_collect_memories_keeps_only_contained_episodes(cls):
    ag = cls(FakeSim(horizon=8), verbose=False, device="cpu")
    ag.collect_memories(4)
    n = len(ag.memory)
    assert ag.logs["init_memories"] == n and n == 4 * 8  # four successful 8-step episodes, nothing else
    r, d = ag.memory._r[:n], ag.memory._d[:n]
    assert int((r == 1000).sum()) == 4 and int(d.sum()) == 4 and int((r == -1000).sum()) == 0
    assert ag.memory.maxlen is None  # `self.memory = deque()`, DQN.py:290


@pytest.mark.parametrize("cls", [A.DQN, A.DQN_BOTH])
def test_learn_writes_reference_style_logs(cls, tmp_path):
    ag = cls(FakeSim(horizon=8), name="unit", verbose=False, device="cpu")
    ag.out_dir = str(tmp_path)
    ag.collect_memories(2)
    ag.learn(6)
    logs_dir = tmp_path / "Logs"
    (name,) = os.listdir(logs_dir)
    assert name.startswith("unit-10s-0k-16m-")
    log = json.load(open(logs_dir / name))
    for key in ("best_reward", "total_rewards", "agent_pos", "agent_deaths", "maps", "init_memories", "total_time",
                "n_episodes", "metadata"):  # DQN.py:23-32, :394
        assert key in log
    assert len(log["total_rewards"]) == 6 and len(log["agent_deaths"]) == 6 and log["n_episodes"] == 6
    assert log["metadata"]["gamma"] == 0.999 and log["metadata"]["width"] == 10
    assert all(tr in (1000 - 7, -1000 - 7) for tr in log["total_rewards"])
    # weights round-trip through the Keras-named .npz
    other = cls(FakeSim(), verbose=False, device="cpu")
    other.load_model(str(tmp_path / "Models" / (name + ".npz")))
    x = torch.rand(3, 10, 10, 3)
    assert torch.allclose(other.model(x), ag.model(x), atol=1e-6)


def test_baseline_run_logs_returns_without_learning(tmp_path):
    ag = A.DQN(FakeSim(horizon=8), name="base", verbose=False, device="cpu")
    ag.out_dir = str(tmp_path)
    ag.collect_memories(10, perform_baseline=True)  # main.py:60-61
    assert len(ag.logs["total_rewards"]) == 10 and len(ag.memory) == 0 and ag.logs["n_episodes"] == 10


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["DQN", "DQN_SARSA", "DQN_DUEL", "DQN_BOTH"])
def test_agents_train_on_the_cuda_environment(name, tmp_path):
    from wildfire_control_python_b200 import ForestFire
    sim = ForestFire(width=10, height=10, seed=31)
    ag = getattr(A, name)(sim, name="gpu", verbose=False)
    ag.out_dir = str(tmp_path)
    assert ag.device.type == "cuda"
    ag.collect_memories(3)
    n = len(ag.memory)
    assert n > 30 and int((ag.memory._r[:n] == 1000).sum()) == 3  # three contained episodes of the walk policy
    ag.learn(4)
    assert len(ag.logs["total_rewards"]) == 4
    (log_name,) = os.listdir(tmp_path / "Logs")
    log = json.load(open(tmp_path / "Logs" / log_name))
    assert len(log["agent_deaths"]) == 4 and log["init_memories"] == n


@pytest.mark.gpu
def test_baseline_policy_through_the_agent_class(tmp_path):
    """`main.py -r -t Baseline`: DQN.collect_memories(E, perform_baseline=True) on the CUDA env."""
    from wildfire_control_python_b200 import ForestFire
    np.random.seed(3)
    sim = ForestFire(width=10, height=10, seed=32)
    ag = A.DQN(sim, name="baseline", verbose=False)
    ag.out_dir = str(tmp_path)
    ag.collect_memories(60, perform_baseline=True)
    rets = np.array(ag.logs["total_rewards"])
    assert len(rets) == 60 and 700 < rets.mean() < 1500  # the thesis reports 1129 over 2500 episodes


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["DQN", "DQN_SARSA"])
def test_batched_demonstration_memories(name):
    """collect_memories over 512 envs at once (walk policy on the device): every stored episode runs from a
    reset to the step that pays the containment bonus, exactly like the reference's per-episode loop."""
    from wildfire_control_python_b200 import ForestFire
    sim = ForestFire(width=14, height=14, seed=33)
    ag = getattr(A, name)(sim, verbose=False)
    steps = ag.collect_memories_batched(200, n_envs=512, k_steps=256, seed=5)
    n = len(ag.memory)
    r, d, s, sp, a = ag.memory._r[:n], ag.memory._d[:n], ag.memory._s[:n], ag.memory._sp[:n], ag.memory._a[:n]
    assert int((r == 1000).sum()) == 200 and ag.logs["init_memories"] == n and steps >= n
    assert 20 * 200 < n < 120 * 200  # the thesis: about 48 transitions per containment at 14x14
    assert int((r == -1000).sum()) == 0 and set(a.unique().tolist()) <= {0, 1, 2, 3}
    ends = torch.nonzero(r == 1000)[:, 0]
    starts = torch.cat([torch.zeros(1, dtype=torch.long, device=ends.device), ends[:-1] + 1])
    # inside an episode the next state of one transition is the state of the following one
    inner = torch.ones(n, dtype=torch.bool, device=ends.device)
    inner[ends] = False
    idx = torch.nonzero(inner)[:, 0]
    assert torch.equal(sp[idx], s[idx + 1])
    # every episode starts from a reset: one burning cell in the centre, agent next to it on a dug cell
    first = s[starts]
    assert bool((first[..., 1].flatten(1).sum(1) == 1).all()) and bool((first[:, 7, 7, 1] == 1).all())
    assert bool((first[..., 0].flatten(1).sum(1) == 1).all()) and bool(((first[..., 2] == 0).flatten(1).sum(1) == 1).all())
    ag.replay()  # and the memory feeds the learner


# ---------------------------------------------------------------------------------------------
# N1 pinned: the update rule against golden vectors restated from the reference's source and Keras 2's code paths
# WITHOUT autograd (oracle/gen_n1_fixture.py: per-sample target loop, hand-written gradients, Keras' clipped Adam).
def _load_n1(name):
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "agents", "n1_replay.npz"))
    return {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "/")}


def _check_n1(name, device):
    cls = getattr(A, name)
    fx = _load_n1(name)
    ag = cls(FakeSim(), verbose=False, device=device)
    keys = sorted(k[3:] for k in fx if k.startswith("w0/"))
    assert keys == sorted(ag.get_weights())  # dense_1..2 (plain) / dense_1..4 (dueling), Keras names
    ag.set_weights({k: fx["w0/" + k] for k in keys})
    with torch.no_grad():  # the target network is a different set of weights
        for lname, m in ag.target.named_children():
            m.weight.copy_(torch.as_tensor(fx[f"target/{lname}/kernel:0"]).t())
            m.bias.copy_(torch.as_tensor(fx[f"target/{lname}/bias:0"]))
    dev = ag.device
    for t in (1, 2):
        b = {k: fx[f"t{t}/batch_{k}"] for k in ("s", "a", "r", "sp", "ap", "d")}
        batch = (torch.as_tensor(b["s"], dtype=torch.uint8, device=dev).view(-1, 10, 10, 3), torch.as_tensor(b["a"], device=dev).long(),
                 torch.as_tensor(b["r"], dtype=torch.float32, device=dev), torch.as_tensor(b["sp"], dtype=torch.uint8, device=dev).view(-1, 10, 10, 3),
                 torch.as_tensor(b["ap"], device=dev).long(), torch.as_tensor(b["d"], device=dev).bool())
        targets = ag.replay_targets(batch).cpu().numpy()
        assert np.allclose(targets, fx[f"t{t}/targets"], rtol=2e-5, atol=2e-3), (name, t, np.abs(targets - fx[f"t{t}/targets"]).max())
        loss = ag.fit_batch(batch)
        assert loss == pytest.approx(float(fx[f"t{t}/loss"]), rel=1e-4)
        got = ag.get_weights()
        for k in keys:
            want = fx[f"t{t}/w/{k}"]
            have = got[k][::16] if got[k].shape[0] == 300 else got[k]
            # one step moves a weight by at most ~lr = 5e-3: 2e-6 absolute is 0.04 % of a step
            assert np.allclose(have, want, rtol=0, atol=2e-6), (name, t, k, np.abs(have - want).max())


@pytest.mark.parametrize("name", ["DQN", "DQN_SARSA", "DQN_DUEL", "DQN_BOTH"])
def test_replay_matches_the_keras_restatement_fixture(name):
    """DQN.replay / DQN_SARSA.replay on DQN.make_network / DQN_DUEL.make_network with Adam(lr, clipvalue=1), two
    consecutive updates: targets, loss and every updated weight against tests/golden/agents/n1_replay.npz (float32 CPU)."""
    _check_n1(name, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["DQN", "DQN_SARSA", "DQN_DUEL", "DQN_BOTH"])
def test_replay_matches_the_keras_restatement_fixture_on_cuda(name):
    torch.backends.cuda.matmul.allow_tf32 = False
    _check_n1(name, "cuda")
