"""Torch restatement of the reference's four learners.

    DQN        DQN.py:9-443          Flatten -> Dense(50, sigmoid) -> Dense(n_actions), Adam(lr=alpha, clipvalue=1), MSE,
                                     replay deque(memory_size), eps-greedy, target network synced every target_update steps
    DQN_SARSA  DQN_SARSA.py:7-191    on-policy target r + gamma * Q_target(s', a')
    DQN_DUEL   DQN_DUEL.py:13-48     dueling head  Q = V + (A - mean A)
    DQN_BOTH   DQN_BOTH.py:4-8       both (same MRO mix-in as the reference)

Same class names, constructor, methods and log format (``Logs/<name>`` JSON readable by the
reference's analyze.py), so ``main.py -r -t DQN|SARSA|DDQN|BOTH|Baseline`` style loops run against the
CUDA environment: ``sim`` is ``wildfire_control_python_b200.ForestFire`` (or anything with the
reference's ``reset / step / render / W.agents / METADATA`` surface).

What is B200-native here: the environment step is the CUDA path behind the C ABI; the replay memory
lives in device tensors and one ``replay()`` is two batched forward passes of the target network and
one optimiser step (the reference issues 2 x batch_size single-sample ``predict`` calls, DQN.py:164-176).
The learning rule itself is unchanged: the regression target of a sample is the TARGET network's
Q-vector with the taken action's entry replaced (DQN.py:166-176), loss = MSE over all outputs.
Keras is not a dependency; weights are saved as ``.npz`` with Keras' layer names
(``dense_1/kernel:0`` ...), kernels in Keras' [in, out] orientation.
"""
from __future__ import annotations

import json
import os
import math
import random
import time

import numpy as np
import torch
from torch import nn

from .constants import DQN_DEFAULTS


class ReplayMemory:
    """``collections.deque(maxlen=...)`` of transitions (DQN.py:20, :205-206), stored as device tensors.

    ``maxlen=None`` grows without bound, like the plain ``deque()`` collect_memories switches to
    (DQN.py:290).  Sampling is uniform without replacement (``random.sample``, DQN.py:161).
    """

    def __init__(self, obs_shape, device, maxlen=None, with_aprime=False):
        self.maxlen, self.device, self.with_aprime = maxlen, device, with_aprime
        self.obs_shape = tuple(obs_shape)
        self._cap = 0
        self._n = 0       # number of valid entries
        self._head = 0    # position of the oldest entry once the ring is full
        self._grow(min(maxlen, 4096) if maxlen else 4096)

    def _grow(self, cap):
        def buf(shape, dtype):
            t = torch.zeros((cap,) + shape, dtype=dtype, device=self.device)
            return t
        new = dict(s=buf(self.obs_shape, torch.uint8), sp=buf(self.obs_shape, torch.uint8), a=buf((), torch.int64),
                   ap=buf((), torch.int64), r=buf((), torch.float32), d=buf((), torch.bool))
        if self._cap:
            for k, t in new.items():
                t[: self._cap] = getattr(self, "_" + k)
        for k, t in new.items():
            setattr(self, "_" + k, t)
        self._cap = cap

    def __len__(self):
        return self._n

    def _as_obs(self, x):
        t = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x)
        return t.reshape(self.obs_shape).to(device=self.device, dtype=torch.uint8)

    def append(self, state, action, reward, sprime, done, aprime=0):
        if self.maxlen is not None and self._n == self.maxlen:
            i = self._head  # full: the oldest entry is dropped, like deque(maxlen)
            self._head = (self._head + 1) % self.maxlen
        else:
            if self._n == self._cap:
                self._grow(min(2 * self._cap, self.maxlen) if self.maxlen else 2 * self._cap)
            i = self._n
            self._n += 1
        self._s[i] = self._as_obs(state)
        self._sp[i] = self._as_obs(sprime)
        self._a[i], self._ap[i], self._r[i], self._d[i] = int(action), int(aprime), float(reward), bool(done)

    def append_bulk(self, states, actions, rewards, sprimes, dones, aprimes=None):
        """Append M transitions held in device tensors (unbounded memories only)."""
        assert self.maxlen is None
        m = int(actions.shape[0])
        while self._n + m > self._cap:
            self._grow(2 * self._cap)
        sl = slice(self._n, self._n + m)
        self._s[sl] = states.reshape((m,) + self.obs_shape).to(self.device, torch.uint8)
        self._sp[sl] = sprimes.reshape((m,) + self.obs_shape).to(self.device, torch.uint8)
        self._a[sl] = actions.to(self.device, torch.int64)
        self._ap[sl] = 0 if aprimes is None else aprimes.to(self.device, torch.int64)
        self._r[sl] = rewards.to(self.device, torch.float32)
        self._d[sl] = dones.to(self.device, torch.bool)
        self._n += m

    def sample(self, batch_size):
        idx = torch.as_tensor(random.sample(range(self._n), batch_size), device=self.device)
        return (self._s[idx], self._a[idx], self._r[idx], self._sp[idx], self._ap[idx], self._d[idx])


class _QNet(nn.Module):
    """DQN.make_network (DQN.py:209-233): Flatten -> Dense(50, sigmoid) -> Dense(n_actions, linear)."""

    def __init__(self, n_in, n_actions):
        super().__init__()
        self.dense_1 = nn.Linear(n_in, 50)
        self.dense_2 = nn.Linear(50, n_actions)
        _keras_init(self)

    def forward(self, x):
        return self.dense_2(torch.sigmoid(self.dense_1(x.flatten(1))))


class _DuelNet(nn.Module):
    """DQN_DUEL.make_network (DQN_DUEL.py:18-48): advantage and value streams, q = v + (a - mean(a))."""

    def __init__(self, n_in, n_actions):
        super().__init__()
        self.dense_1 = nn.Linear(n_in, 50)       # advantage stream
        self.dense_2 = nn.Linear(50, n_actions)
        self.dense_3 = nn.Linear(n_in, 50)       # value stream
        self.dense_4 = nn.Linear(50, 1)
        _keras_init(self)

    def forward(self, x):
        x = x.flatten(1)
        adv = self.dense_2(torch.sigmoid(self.dense_1(x)))
        val = self.dense_4(torch.sigmoid(self.dense_3(x)))
        return val + (adv - adv.mean(dim=1, keepdim=True))


class KerasAdam(torch.optim.Optimizer):
    """``keras.optimizers.Adam(lr, clipvalue)`` of Keras 2 (the reference's ``Adam(lr=self.alpha, clipvalue=1)``,
    DQN.py:227-230), update for update: every gradient element is clipped to [-clipvalue, clipvalue], then
    ``lr_t = lr * sqrt(1 - beta_2^t) / (1 - beta_1^t)``, ``p -= lr_t * m / (sqrt(v) + epsilon)`` with
    ``epsilon = K.epsilon() = 1e-7`` added to the UNCORRECTED second moment (``torch.optim.Adam`` adds its epsilon after
    the bias correction, i.e. an effective epsilon sqrt(1 - beta_2^t) times smaller).  Pinned by tests/golden/agents/n1_replay.npz."""

    def __init__(self, params, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, clipvalue=None):
        super().__init__(params, dict(lr=lr, beta_1=beta_1, beta_2=beta_2, epsilon=epsilon, clipvalue=clipvalue))

    @torch.no_grad()
    def step(self):
        for group in self.param_groups:
            b1, b2 = group["beta_1"], group["beta_2"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["t"], st["m"], st["v"] = 0, torch.zeros_like(p), torch.zeros_like(p)
                g = p.grad if group["clipvalue"] is None else p.grad.clamp(-group["clipvalue"], group["clipvalue"])
                st["t"] += 1
                st["m"].mul_(b1).add_(g, alpha=1.0 - b1)
                st["v"].mul_(b2).addcmul_(g, g, value=1.0 - b2)
                lr_t = group["lr"] * math.sqrt(1.0 - b2 ** st["t"]) / (1.0 - b1 ** st["t"])
                p.addcdiv_(st["m"], st["v"].sqrt().add_(group["epsilon"]), value=-lr_t)


def _keras_init(net):
    """Keras Dense defaults: glorot_uniform kernel, zero bias."""
    for m in net.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            nn.init.zeros_(m.bias)


class DQN:
    def __init__(self, sim, name="no_name", verbose=True, device=None):
        # Constants and such (DQN.py:10-17)
        self.sim = sim
        self.name = name
        self.METADATA = dict(DQN_DEFAULTS)
        self.METADATA.update(sim.METADATA)
        self.action_size = self.sim.n_actions
        self.DEBUG = getattr(sim, "DEBUG", 1)
        self.verbose = verbose
        if device is None:
            device = getattr(getattr(sim, "_batch", None), "device", None) or ("cuda" if torch.cuda.is_available() else "cpu")
        self.device = torch.device(device)
        self.obs_shape = (self.sim.W.WIDTH, self.sim.W.HEIGHT, self.sim.W.DEPTH)

        # DQN memory (DQN.py:20)
        self.memory = self._new_memory(self.METADATA["memory_size"])

        # Information to save to file (DQN.py:23-32)
        self.logs = {
            "best_reward": -10000,
            "total_rewards": list(),
            "agent_pos": list(),
            "agent_deaths": list(),
            "maps": list(),
            "init_memories": 0,
            "total_time": 0,
            "n_episodes": 0,
        }

        # DQN parameters (DQN.py:35-41)
        self.max_eps = self.METADATA["max_eps"]
        self.min_eps = self.METADATA["min_eps"]
        self.eps_decay_rate = self.METADATA["eps_decay_rate"]
        self.eps = self.max_eps
        self.gamma = self.METADATA["gamma"]
        self.alpha = self.METADATA["alpha"]
        self.target_update_freq = self.METADATA["target_update"]

        # Network and target network (DQN.py:44-46)
        self.model = self.make_network()
        self.target = self.make_network()
        self.target.load_state_dict(self.model.state_dict())
        # Adam(lr=alpha, clipvalue=1), Keras' update rule; loss = mse (DQN.py:227-230)
        self.optimizer = KerasAdam(self.model.parameters(), lr=self.alpha, clipvalue=1.0)

        if self.verbose:
            width, height = self.METADATA["width"], self.METADATA["height"]
            print("\n\t[Parameters]")
            print("[decay]", self.METADATA["eps_decay_rate"])
            print("[alpha]", self.METADATA["alpha"])
            print("[gamma]", self.METADATA["gamma"])
            print("[batch]", self.METADATA["batch_size"])
            print("[size]", f"{width}x{height}")
            print("[wind speed]", self.METADATA["wind"][0] if self.METADATA["wind"] != "random" else "random")
            print("[target upd]", self.METADATA["target_update"], "\n")

    # ------------------------------------------------------------------ learning
    _SARSA = False

    def _new_memory(self, maxlen):
        return ReplayMemory(self.obs_shape, self.device, maxlen=maxlen, with_aprime=self._SARSA)

    def _as_batch(self, state):
        """np.reshape(state, [1] + shape) of the reference; a device float tensor here."""
        t = state if torch.is_tensor(state) else torch.as_tensor(np.asarray(state))
        return t.reshape((1,) + self.obs_shape).to(device=self.device, dtype=torch.float32)

    def learn(self, n_episodes=1000):
        """The training loop of DQN.learn (DQN.py:65-153) and DQN_SARSA.learn (DQN_SARSA.py:12-100): one episode
        after another, one ``replay`` per step once the memory holds more than a batch, target network
        refreshed every ``target_update`` steps, epsilon decayed per episode, logs written at the end.
        The on-policy variant picks the next action BEFORE storing the transition (it is part of it)."""
        t_run = time.time()
        self.logs["n_episodes"] = n_episodes
        until_sync = self.target_update_freq
        batch = self.METADATA["batch_size"]
        for episode in range(n_episodes):
            t_episode = time.time()
            state = self._as_batch(self.sim.reset())
            if self.DEBUG > 0 and not self._SARSA:  # only DQN.learn records the start cell (DQN.py:89-91)
                me = self.sim.W.agents[0]
                self.logs["agent_pos"].append((me.x, me.y))
            action = self.choose_action(state) if self._SARSA else None
            episode_return, done = 0, False
            while not done:
                if not self._SARSA:
                    action = self.choose_action(state)
                observed, reward, done, _ = self.sim.step(action)
                observed = self._as_batch(observed)
                if self._SARSA:
                    upcoming = self.choose_action(observed)
                    self.remember(state, action, reward, observed, upcoming, done)
                else:
                    upcoming = None
                    self.remember(state, action, reward, observed, done)
                if len(self.memory) > batch:
                    self.replay()
                until_sync -= 1
                if until_sync == 0:
                    until_sync = self.target_update_freq
                    self.target.load_state_dict(self.model.state_dict())
                state, action = observed, upcoming
                episode_return += reward
            self._end_of_episode(episode, episode_return, t_episode)
        self.logs["total_time"] = round(time.time() - t_run, 3)
        self.write_data()

    def _end_of_episode(self, episode, total_reward, t0):  # DQN.py:120-148
        dead = len(self.sim.W.agents) == 0
        if self.DEBUG > 0:
            self.logs["agent_deaths"].append(dead)
        if total_reward >= 0.9 * self.logs["best_reward"] or total_reward > 300:
            map_string = self._render_string()
            if total_reward > self.logs["best_reward"]:
                self.logs["best_reward"] = total_reward
            if self.DEBUG > 0:
                self.logs["maps"].append([episode, map_string])
        if self.verbose:
            print(f"[Episode {episode + 1}]\tTime: {round(time.time() - t0, 3)}")
            print(f"\t\tEpsilon: {round(self.eps, 3)}")
            print(f"\t\tAgent dead: {dead}")
            print(f"\t\tReward: {round(total_reward, 0)}\n")
        self.decay_epsilon(episode)
        self.logs["total_rewards"].append(total_reward)

    def _render_string(self):
        try:
            return self.sim.render(print_map=self.verbose)
        except TypeError:
            return self.sim.render()

    def _bootstrap(self, q_next, aprime):
        """Value of S' used in the target: max_a Q_target(S', a) (DQN.py:174-175)."""
        return q_next.max(dim=1).values

    def replay_targets(self, batch):
        """The regression targets of one replay batch (DQN.py:164-180): the target network's Q-vector
        of S with the taken action's entry replaced by r (terminal) or r + gamma * bootstrap(S')."""
        s, a, r, sp, ap, d = batch
        with torch.no_grad():
            pred = self.target(s.float())
            boot = self._bootstrap(self.target(sp.float()), ap)
            upd = torch.where(d, r, r + self.gamma * boot)
            pred[torch.arange(len(a), device=pred.device), a] = upd
        return pred

    def replay(self):  # DQN.py:156-185
        return self.fit_batch(self.memory.sample(self.METADATA["batch_size"]))

    def fit_batch(self, batch):
        """``self.model.fit(states, predictions, epochs=1)`` on one batch of transitions (DQN.py:181-185): with the
        reference's batch_size = 32 = Keras' default mini-batch that is exactly one clipped-Adam update on the MSE."""
        targets = self.replay_targets(batch)
        loss = nn.functional.mse_loss(self.model(batch[0].float()), targets)
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        self.optimizer.step()  # clips every gradient element to [-1, 1] first (clipvalue=1)
        return float(loss.detach())

    def choose_action(self, state, eps=None):  # DQN.py:188-196
        eps_threshold = self.eps if eps is None else eps
        if random.uniform(0, 1) > eps_threshold:
            with torch.no_grad():
                return int(self.model(self._as_batch(state)).argmax(dim=1)[0])
        return int(np.random.choice(self.METADATA["n_actions"]))

    def decay_epsilon(self, episode_num=None):  # DQN.py:199-202
        self.eps = self.min_eps + (self.max_eps - self.min_eps) * np.exp(-self.eps_decay_rate * episode_num)

    def remember(self, state, action, reward, sprime, done):  # DQN.py:205-206
        self.memory.append(state, action, reward, sprime, done)

    def make_network(self):  # DQN.py:209-233
        return _QNet(int(np.prod(self.obs_shape)), self.action_size).to(self.device)

    # ------------------------------------------------------------------ helpers
    def play_optimal(self, eps=0, delay=0.1):  # DQN.py:240-253
        done = False
        total_reward = 0
        state = self.sim.reset()
        while not done:
            self.sim.render()
            self.show_info(self._as_batch(state))
            action = self.choose_action(state, eps=eps)
            state, reward, done, _ = self.sim.step(action)
            total_reward += reward
            time.sleep(delay)
        self.sim.render()
        print(f"Total reward: {total_reward}")
        return total_reward

    def show_info(self, state):  # DQN.py:256-276
        print(f"Wind Speed: {self.sim.W.wind_speed}")
        print(f"Wind direction: {self.sim.W.wind_vector}")
        with torch.no_grad():
            qvals = self.model(self._as_batch(state))[0].cpu().numpy()
        key_map = {0: "N", 1: "S", 2: "E", 3: "W", 4: "D", 5: " "}
        print("| ", end="")
        for idx, val in enumerate(qvals):
            val = round(float(val), 2)
            extra_space = " " if val > 0 else ""
            print(f"{key_map[idx]} : {extra_space}{val:.2f} | ", end="")
            if idx == 1:
                print("\n| ", end="")
        print(f"\nBest Action: {key_map[int(np.argmax(qvals))]}\n")

    def collect_memories(self, num_of_episodes=100, perform_baseline=False):
        """Demonstration data (DQN.py:286-348, DQN_SARSA.py:148-191): the heuristic walk round the fire drives the
        env; an episode's transitions enter the (now unbounded, DQN.py:290) memory only if it pays the
        containment bonus, and it is cut there; stop after ``num_of_episodes`` such episodes.
        ``perform_baseline`` (main.py:60-61): store nothing, play ``num_of_episodes`` whole episodes of the
        heuristic policy and log their returns and deaths."""
        if not num_of_episodes:
            return
        self.memory = self._new_memory(None)
        bonus = self.METADATA["contained_bonus"]
        kept, played = 0, 0
        while True:
            episode, episode_return, done = [], 0, False
            state = self._as_batch(self.sim.reset())
            action = self.choose_randomwalk_action() if self._SARSA else None
            while not done:
                if not self._SARSA:
                    action = self.choose_randomwalk_action()
                observed, reward, done, _ = self.sim.step(action)
                observed = self._as_batch(observed)
                if self._SARSA:
                    upcoming = self.choose_randomwalk_action()
                    episode.append((state, action, reward, observed, upcoming, done))
                else:
                    upcoming = None
                    episode.append((state, action, reward, observed, done))
                state, action = observed, upcoming
                episode_return += reward
                if not perform_baseline and reward == bonus:
                    kept += 1
                    for transition in episode:
                        self.remember(*transition)
                    if kept == num_of_episodes:
                        self.logs["init_memories"] = len(self.memory)
                        return
                    break
            if perform_baseline:
                self.logs["total_rewards"].append(episode_return)
                self.logs["agent_deaths"].append(len(self.sim.W.agents) == 0)
                if self.verbose and played % 100 == 0:
                    print(f"Episode {played}/{num_of_episodes}")
                played += 1
                if played == num_of_episodes:
                    self.logs["n_episodes"] = num_of_episodes
                    break
        self.write_data()

    def collect_memories_batched(self, num_of_episodes=100, n_envs=1024, k_steps=256, seed=0):
        """``collect_memories`` (DQN.py:286-348, DQN_SARSA.py:148-191) over ``n_envs`` environments at once:
        the heuristic walk policy runs ON THE DEVICE (``wf_rollout_policy``), whole rollouts of ``k_steps`` steps
        per launch, and the transitions of the episodes that contain the fire (from their reset up to and
        including the step that pays the containment bonus) are gathered into the replay memory.
        Only episodes that start in the first half of a rollout are considered, so that slow containments
        are not under-represented; an episode still undecided at the end of its rollout is dropped.
        Returns the number of environment steps simulated."""
        from .batched import BatchedForestFire
        from .constants import DQN_DEFAULTS as _D
        if not num_of_episodes:
            return 0
        keys = {k: v for k, v in self.METADATA.items() if k not in _D and k not in ("debug", "a_speed_iter", "seed", "auto_reset")}
        env = BatchedForestFire(n_envs, device=self.device if self.device.type == "cuda" else None, auto_reset=True, seed=seed, **keys)
        bonus = self.METADATA["contained_bonus"]
        self.memory = self._new_memory(None)
        success, simulated = 0, 0
        try:
            while success < num_of_episodes:
                obs0 = env.reset().clone()
                obs, rew, done, acts = env.rollout(k_steps, policy="walk", return_actions=True)
                simulated += k_steps * n_envs
                rew_c, done_c = rew.cpu().numpy(), done.cpu().numpy()
                t_idx, e_idx = [], []
                for i in range(n_envs):
                    bounds = list(np.nonzero(done_c[:, i])[0]) + [k_steps - 1]
                    start = 0
                    for end in bounds:
                        if start >= k_steps // 2 or start > end:
                            break
                        hit = np.nonzero(rew_c[start:end + 1, i] == bonus)[0]
                        if len(hit):
                            t_idx.append(np.arange(start, start + hit[0] + 1))
                            e_idx.append(np.full(hit[0] + 1, i))
                            success += 1
                            if success == num_of_episodes:
                                break
                        start = end + 1
                    if success == num_of_episodes:
                        break
                if t_idx:
                    t = torch.as_tensor(np.concatenate(t_idx), device=obs.device)
                    e = torch.as_tensor(np.concatenate(e_idx), device=obs.device)
                    prev = torch.where((t > 0)[:, None, None, None], obs[(t - 1).clamp(min=0), e], obs0[e])
                    nxt = acts[(t + 1).clamp(max=k_steps - 1), e]
                    self.memory.append_bulk(prev, acts[t, e], rew[t, e], obs[t, e], done[t, e], nxt if self._SARSA else None)
        finally:
            env.close()
        self.logs["init_memories"] = len(self.memory)
        return simulated

    @staticmethod
    def _clockwise_options(ax, ay, mid_x, mid_y):
        """The two moves that keep the agent circling the fire origin clockwise (DQN.py:369-376); actions are
        0 N, 1 S, 2 E, 3 W.  The four half-open quadrants tile the map except the origin itself."""
        if ax >= mid_x and ay > mid_y:
            return (1, 3)
        if ax > mid_x and ay <= mid_y:
            return (1, 2)
        if ax <= mid_x and ay < mid_y:
            return (0, 2)
        if ax < mid_x and ay >= mid_y:
            return (0, 3)
        raise UnboundLocalError("the agent stands on the fire origin: the reference has no move for that cell")

    def choose_randomwalk_action(self, avoid_fire=True):
        """DQN.choose_randomwalk_action (DQN.py:353-389): a random one of the two clockwise moves, redrawn (at most
        11 times) while it would step onto a burning cell.  0 when the agent is gone (SARSA asks anyway)."""
        live = self.sim.W.agents
        if not live:
            return 0
        me = live[0]
        options = self._clockwise_options(me.x, me.y, int(self.sim.W.WIDTH / 2), int(self.sim.W.HEIGHT / 2))
        for attempt in range(12):
            action = options[int(np.random.choice(2))]
            if not avoid_fire or not me.fire_in_direction(action):
                break
        return action

    # ------------------------------------------------------------------ persistence
    out_dir = "."  # Logs/ and Models/ are created below this directory (the reference uses the cwd)

    def write_data(self):  # DQN.py:392-424
        self.logs["metadata"] = self.METADATA
        n_episodes = self.logs["n_episodes"]
        n_episodes = n_episodes / 1000 if n_episodes >= 1000 else 0
        memories = self.logs["init_memories"]
        name = self.sim.get_name(self.sim.W.WIDTH, int(n_episodes), memories, self.name)
        logs_dir, models_dir = os.path.join(self.out_dir, "Logs"), os.path.join(self.out_dir, "Models")
        counter = 0
        while os.path.isfile(os.path.join(logs_dir, name)) or os.path.isfile(os.path.join(models_dir, name + ".npz")):
            if counter > 0:
                name = name[: -len(str(counter))]
            name = name + str(counter)
            counter += 1
        os.makedirs(logs_dir, exist_ok=True)
        os.makedirs(models_dir, exist_ok=True)
        self.save_model(name)
        with open(os.path.join(logs_dir, name), "w") as file:
            json.dump(self.logs, file, default=_jsonable)
        return name

    def get_weights(self):
        """Keras-style weight dict: ``dense_k/kernel:0`` [in, out] and ``dense_k/bias:0``."""
        out = {}
        for lname, m in self.model.named_children():
            out[f"{lname}/kernel:0"] = m.weight.detach().t().contiguous().cpu().numpy()
            out[f"{lname}/bias:0"] = m.bias.detach().cpu().numpy()
        return out

    def set_weights(self, weights):
        with torch.no_grad():
            for lname, m in self.model.named_children():
                m.weight.copy_(torch.as_tensor(np.asarray(weights[f"{lname}/kernel:0"])).t())
                m.bias.copy_(torch.as_tensor(np.asarray(weights[f"{lname}/bias:0"])))
        self.target.load_state_dict(self.model.state_dict())

    def load_keras_weights(self, path):
        """Load a network trained by the REFERENCE: a Keras ``save_weights`` HDF5 file from its ``Models/``
        tree (DQN.py:441-443), read without h5py/Keras by ``keras_h5.read_keras_weights``."""
        from .keras_h5 import canonical_dense_names, read_keras_weights
        self.set_weights(canonical_dense_names(read_keras_weights(path)))

    def evaluate_batched(self, n_envs=4096, episodes_per_env=1, eps=None, seed=0, max_steps=100000, in_kernel=None):
        """Roll the current policy out on ``n_envs`` CUDA environments at once (eps-greedy, DQN.py:188-196) and
        return (total reward, agent died) of the first ``episodes_per_env`` episodes of every env -- the
        quantity ``logs['total_rewards']`` / ``logs['agent_deaths']`` hold per training episode.

        ``in_kernel`` (default: whenever the grid fits the warp kernel family, <= 32x32): the Q-network is
        evaluated INSIDE the step kernel (``wf_rollout_policy(WF_POLICY_MLP)``), 64 steps per launch;
        otherwise torch evaluates it between two ``wf_step`` calls."""
        from .batched import BatchedForestFire
        from .constants import DQN_DEFAULTS as _D
        eps = self.min_eps if eps is None else eps
        keys = {k: v for k, v in self.METADATA.items() if k not in _D and k not in ("debug", "a_speed_iter", "seed", "auto_reset")}
        env = BatchedForestFire(n_envs, device=self.device, auto_reset=True, seed=seed, **keys)
        if in_kernel is None:
            in_kernel = env.kernel_family == "warp"
        if in_kernel:
            try:
                return self._evaluate_in_kernel(env, n_envs, episodes_per_env, eps, max_steps)
            finally:
                env.close()
        dev = self.device
        gen = torch.Generator(device=dev).manual_seed(seed)
        returns = torch.zeros((n_envs, episodes_per_env), dtype=torch.float64, device=dev)
        died = torch.zeros((n_envs, episodes_per_env), dtype=torch.bool, device=dev)
        acc = torch.zeros(n_envs, dtype=torch.float64, device=dev)
        n_done = torch.zeros(n_envs, dtype=torch.int64, device=dev)
        rows = torch.arange(n_envs, device=dev)
        death_penalty = float(self.METADATA["death_penalty"])
        try:
            obs = env.reset()
            for step in range(max_steps):
                with torch.no_grad():
                    greedy = self.model(obs.float()).argmax(dim=1).to(torch.int32)
                explore = torch.rand(n_envs, device=dev, generator=gen) < eps
                rnd = torch.randint(0, self.action_size, (n_envs,), device=dev, dtype=torch.int32, generator=gen)
                obs, rew, done, _ = env.step(torch.where(explore, rnd, greedy))
                acc += rew
                live = done & (n_done < episodes_per_env)
                slot = n_done.clamp(max=episodes_per_env - 1)
                returns[rows[live], slot[live]] = acc[live]
                died[rows[live], slot[live]] = rew[live] == death_penalty
                acc = torch.where(done, torch.zeros_like(acc), acc)
                n_done += done.to(torch.int64)
                if step % 16 == 15 and bool((n_done >= episodes_per_env).all()):
                    break
            else:
                raise RuntimeError("evaluate_batched: some episodes did not finish")
        finally:
            env.close()
        return returns.flatten().cpu().numpy(), died.flatten().cpu().numpy()

    def _evaluate_in_kernel(self, env, n_envs, episodes_per_env, eps, max_steps, chunk=64):
        w = self.get_weights()  # dense_1 / dense_2 are the advantage stream of a dueling head: argmax Q = argmax A
        env.set_policy_mlp(w["dense_1/kernel:0"], w["dense_1/bias:0"], w["dense_2/kernel:0"], w["dense_2/bias:0"], eps=eps)
        env.reset()
        death_penalty = float(self.METADATA["death_penalty"])
        returns = np.zeros((n_envs, episodes_per_env))
        died = np.zeros((n_envs, episodes_per_env), dtype=bool)
        n_done = np.zeros(n_envs, dtype=np.int64)
        carry = np.zeros(n_envs)  # reward of the running episode before this chunk
        for _ in range(0, max_steps, chunk):
            _, rew, done = env.rollout(chunk, policy="mlp", obs=False)
            rew, done = rew.cpu().numpy(), done.cpu().numpy()
            c = carry[None, :] + np.cumsum(rew, axis=0)
            e_idx, k_idx = np.nonzero(done.T)  # episode ends, sorted by env then step
            at_end = c[k_idx, e_idx]
            same = np.r_[False, e_idx[1:] == e_idx[:-1]]
            ret = at_end - np.where(same, np.r_[0.0, at_end[:-1]], 0.0)
            first = np.r_[True, ~same[1:]] if len(e_idx) else np.zeros(0, bool)
            rank = np.arange(len(e_idx)) - np.maximum.accumulate(np.where(first, np.arange(len(e_idx)), 0))
            slot = n_done[e_idx] + rank
            keep = slot < episodes_per_env
            returns[e_idx[keep], slot[keep]] = ret[keep]
            died[e_idx[keep], slot[keep]] = rew[k_idx[keep], e_idx[keep]] == death_penalty
            last = np.zeros(n_envs)
            if len(e_idx):
                is_last = np.r_[e_idx[1:] != e_idx[:-1], True]
                last[e_idx[is_last]] = at_end[is_last]
            carry = c[-1] - last
            n_done += np.bincount(e_idx, minlength=n_envs)
            if (n_done >= episodes_per_env).all():
                return returns.flatten(), died.flatten()
        raise RuntimeError("evaluate_batched: some episodes did not finish")

    def save_model(self, name):  # DQN.py:441-443
        path = os.path.join(self.out_dir, "Models", name)
        np.savez(path, **self.get_weights())
        return path + ".npz"

    def load_model(self, path=None):  # DQN.py:427-438 (interactive when no path is given)
        if path is None:
            all_names = sorted(os.listdir(os.path.join(self.out_dir, "Models")))
            print("\nChoose a model from the list below to load:")
            for idx, n in enumerate(all_names):
                print(f"\t[{idx}] {n}")
            try:
                path = os.path.join(self.out_dir, "Models", all_names[int(input(f"Select one [0-{len(all_names) - 1}]: \n"))])
            except (ValueError, IndexError):
                print("Invalid Selection")
                return
        with np.load(path) as z:
            self.set_weights({k: z[k] for k in z.files})
        if self.verbose:
            print("Model loaded!")


def _jsonable(o):
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, np.ndarray):
        return o.tolist()
    if isinstance(o, (np.bool_,)):
        return bool(o)
    raise TypeError(f"not JSON serialisable: {type(o)}")


class DQN_SARSA(DQN):
    _SARSA = True

    def __init__(self, sim, name="no_name", verbose=True, device=None):
        DQN.__init__(self, sim, name, verbose, device)

    def _bootstrap(self, q_next, aprime):
        """On-policy value of S': Q_target(S', A') (DQN_SARSA.py:120-121)."""
        return q_next.gather(1, aprime[:, None])[:, 0]

    def remember(self, state, action, reward, sprime, aprime, done):  # DQN_SARSA.py:134-135
        self.memory.append(state, action, reward, sprime, done, aprime)

    def collect_memories(self, num_of_successes=100):  # DQN_SARSA.py:148-191 (no baseline mode there)
        return DQN.collect_memories(self, num_of_successes, perform_baseline=False)


class DQN_DUEL(DQN):
    def __init__(self, sim, name="no_name", verbose=True, device=None):
        DQN.__init__(self, sim, name, verbose, device)

    def make_network(self):  # DQN_DUEL.py:18-48
        return _DuelNet(int(np.prod(self.obs_shape)), self.action_size).to(self.device)


class DQN_BOTH(DQN_SARSA, DQN_DUEL):  # DQN_BOTH.py:4-8
    def __init__(self, sim, name="no_name", verbose=True, device=None):
        DQN_SARSA.__init__(self, sim, name, verbose, device)
