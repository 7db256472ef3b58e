"""ForestFire -- single-environment facade with the reference's attribute surface.

Mirrors ``Simulation.forest_fire.ForestFire`` (forest_fire.py:18-106) so that the reference's
agents (DQN.py, DQN_SARSA.py, DQN_DUEL.py, DQN_BOTH.py) and ``misc.py`` can drive the CUDA
environment unmodified.  What they touch:

    sim.reset() -> ndarray (W, H, 3) float64          sim.step(a) -> [state, reward, done, {}]
    sim.render() -> str                               sim.METADATA, sim.n_actions, sim.DEBUG
    sim.W.WIDTH / HEIGHT / DEPTH                      sim.W.agents  ([] once dead; [0].x/.y,
    sim.W.wind_speed / wind_vector / RUNNING            [0].fire_in_direction(a), .dead, .digging)
    sim.W.env  (W, H, 9) float64 view, read-only      sim.layer, sim.get_name, sim.width/height

It is a thin view over a ``BatchedForestFire`` of one env (or over env ``index`` of an existing
batch); every call goes through the C ABI -- there is no Python re-implementation of the step.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import _lib
from .batched import BatchedForestFire
from .constants import ascii_of_type, layer, types


def get_name(size, episodes, memories, name):  # Simulation/utility.py:81-85
    return f"""{name}-{size}s-{episodes}k-{memories}m-{time.strftime("%m-%d-%H%M")}"""


_ACTIONS = {"N": 0, "S": 1, "E": 2, "W": 3, "D": 4}


class _AgentView:
    """``World.agents[0]`` (environment.py:99-171), read-only."""

    def __init__(self, world):
        self._w = world

    @property
    def x(self):
        return int(self._w._scalars()[_lib.S_AX])

    @property
    def y(self):
        return int(self._w._scalars()[_lib.S_AY])

    @property
    def dead(self):
        return bool(self._w._scalars()[_lib.S_DEAD])

    @property
    def digging(self):
        return bool(self._w._scalars()[_lib.S_DIGGING])

    def is_dead(self):
        return self.dead or self._w.is_burning((self.x, self.y))

    def _direction_to_coords(self, direction):  # environment.py:163-171
        d = _ACTIONS.get(direction, direction)
        return {0: (self.x, self.y - 1), 1: (self.x, self.y + 1), 2: (self.x + 1, self.y), 3: (self.x - 1, self.y)}[d]

    def fire_in_direction(self, direction):  # environment.py:158-160
        nx, ny = self._direction_to_coords(direction)
        return bool(self._w.inbounds(nx, ny) and self._w.is_burning((nx, ny)))


class _WorldView:
    """``ForestFire.W`` (environment.py:174-402), read-only views of the device state."""

    def __init__(self, sim):
        self._sim = sim
        self.WIDTH, self.HEIGHT, self.DEPTH = sim.width, sim.height, 3
        self._agent = _AgentView(self)

    def _state(self):
        return self._sim._state()

    def _scalars(self):
        return self._state()["scalars"]

    @property
    def agents(self):
        return [self._agent] if self._scalars()[_lib.S_ALIVE] else []

    @property
    def RUNNING(self):
        return bool(self._scalars()[_lib.S_RUNNING])

    @property
    def fire_at_border(self):
        return bool(self._scalars()[_lib.S_FIRE_AT_BORDER])

    @property
    def wind_speed(self):
        return float(self._sim._batch.wind_speed_table[self._scalars()[_lib.S_WIND_ID]])

    @property
    def wind_vector(self):
        s = self._scalars()
        return (int(s[_lib.S_WIND_X]), int(s[_lib.S_WIND_Y]))

    @property
    def burning_cells(self):
        xs, ys = np.nonzero(self._state()["burning"])
        return {(int(x), int(y)) for x, y in zip(xs, ys)}

    @property
    def border_points(self):
        """Non-empty until the containment bonus has been paid (environment.py:345-377, quirk Q4)."""
        if self._scalars()[_lib.S_LATCHED]:
            return []
        W, H = self.WIDTH, self.HEIGHT
        pts = []
        for x in range(W):
            pts += [[x, 0], [x, H - 1]]
        for y in range(H):
            pts += [[0, y], [H - 1, y]]
        return pts

    def inbounds(self, x, y):  # environment.py:225-226
        return 0 <= x < self.WIDTH and 0 <= y < self.HEIGHT

    def traversable(self, x, y):
        return bool(self._state()["type"][x, y] != types["water"])

    def is_burning(self, cell):  # environment.py:249-251 -- a type test
        return bool(self._state()["type"][cell[0], cell[1]] == types["fire"])

    def is_burnable(self, cell):
        return bool(self._state()["type"][cell[0], cell[1]] == types["grass"])

    def get_state(self):  # environment.py:399-402
        return self._sim._obs_f64()

    @property
    def env(self):
        """The reference's (W, H, 9) float64 array (environment.py:38-50), rebuilt on demand.
        ``gray`` carries the type code instead of a grey level (render() does not need it)."""
        st = self._state()
        W, H = self.WIDTH, self.HEIGHT
        m = self._sim.METADATA
        out = np.zeros((W, H, 9))
        out[:, :, layer["type"]] = st["type"]
        out[:, :, layer["gray"]] = st["type"]
        out[:, :, layer["temp"]] = st["temp"]
        out[:, :, layer["heat"]] = m["heat"]
        out[:, :, layer["fuel"]] = st["fuel"]
        out[:, :, layer["threshold"]] = m["threshold"]
        out[:, :, layer["agent_pos"]] = st["apos"]
        out[:, :, layer["fire_mobility"]] = np.where(st["fm_inf"] != 0, np.inf, 1.0)
        out[:, :, layer["agent_mobility"]] = np.where(st["type"] == types["water"], np.inf, 1.0)
        return out


class ForestFire:
    """Drop-in for ``Simulation.forest_fire.ForestFire``; keyword arguments replace the reference's
    edit-the-global-METADATA configuration (same key names)."""

    def __init__(self, batch: BatchedForestFire = None, index: int = 0, **metadata):
        self._batch = batch if batch is not None else BatchedForestFire(1, **metadata)
        if self._batch.n_envs != 1 or index != 0:
            # step() acts on the whole batch: a no-op action still ages the other envs (t, the shared a_speed_iter, the
            # fire tick, auto-reset), so a view into a larger batch would change its neighbours' trajectories.
            raise ValueError("ForestFire is a single-env facade: pass a BatchedForestFire with n_envs == 1 "
                             "(step whole batches through BatchedForestFire itself)")
        self._i = 0
        self.METADATA = self._batch.METADATA
        self.DEBUG = self.METADATA.get("debug", 1)
        self.layer = layer
        self.get_name = get_name
        self.width, self.height = self._batch.width, self._batch.height
        self.n_actions = self._batch.n_actions
        self._cache = None
        self._actions = torch.full((self._batch.n_envs,), -1, dtype=torch.int32, device=self._batch.device)
        self.W = _WorldView(self)
        self.reset()  # World.__init__ resets once (environment.py:183)

    # -- helpers ---------------------------------------------------------------------------------
    def _state(self):
        if self._cache is None:
            st = self._batch.get_state()
            self._cache = {k: v[self._i].cpu().numpy() for k, v in st.items()}
        return self._cache

    def _obs_f64(self):
        return self._batch.observe()[self._i].cpu().numpy().astype(np.float64)

    # -- the reference API -------------------------------------------------------------------------
    def reset(self):
        self._cache = None
        if self._batch.n_envs == 1:
            obs = self._batch.reset()
        else:
            mask = torch.zeros(self._batch.n_envs, dtype=torch.uint8, device=self._batch.device)
            mask[self._i] = 1
            obs = self._batch.reset(mask=mask)
        return obs[self._i].cpu().numpy().astype(np.float64)

    def step(self, action):
        a = _ACTIONS.get(action, action) if isinstance(action, str) else int(action)
        if action == "D" and not self.METADATA["allow_dig_toggle"]:
            a = -1
        if not isinstance(a, int):
            a = -1
        sc = self._state()["scalars"]
        if (0 <= a < 4 or (a == 4 and self.METADATA["allow_dig_toggle"])) and not sc[_lib.S_ALIVE]:
            raise IndexError("list index out of range")  # the reference's agents[0] on an empty list (Q8)
        if not sc[_lib.S_RUNNING]:
            raise RuntimeError("step() on a finished episode: call reset() (the batched env freezes finished envs)")
        self._cache = None
        self._actions.fill_(-1)
        self._actions[self._i] = a
        obs, rew, done, info = self._batch.step(self._actions)
        r = float(rew[self._i])
        for key in ("contained_bonus", "death_penalty", "default_reward"):  # the reference returns these ints as-is
            if r == self.METADATA[key]:
                r = self.METADATA[key]
                break
        return [obs[self._i].cpu().numpy().astype(np.float64), r, bool(done[self._i]), {}]

    def render(self, print_map: bool = True):
        """ASCII rendering (forest_fire.py:57-82): '+' grass '@' fire '#' burnt '0' dirt 'x' water 'A' agent."""
        st = self._state()
        sc = st["scalars"]
        rows = []
        for y in range(self.height):
            row = ""
            for x in range(self.width):
                if sc[_lib.S_ALIVE] and (sc[_lib.S_AX], sc[_lib.S_AY]) == (x, y):
                    row += "A"
                else:
                    row += ascii_of_type[int(st["type"][x, y])]
            rows.append(row)
        if print_map:
            print(" " + "".join(str(x % 10) for x in range(self.width)))
            for y, row in enumerate(rows):
                print(str(y % 10) + row)
            print("")
        return "\n" + "\n".join(rows) + "\n"

    def update(self):
        raise NotImplementedError("the fire tick runs inside step(); it is not exposed separately")
