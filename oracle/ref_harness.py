"""Run the UNMODIFIED reference environment under the shared Philox stream.

TEST INFRASTRUCTURE.  Only usable in the build container, where the reference
tree is mounted read-only at /root/reference (it does not exist on the GPU box).
Used by ``oracle/gen_golden.py`` to write the fixtures under ``tests/golden/``
and by ``oracle/validate_oracle.py`` to pin the C restatement.

What is shimmed (nothing of the reference is copied or edited):
  * ``pyastar``: the reference binding (pyastar/pyastar.py:9-22) loads ``astar.so``
    from the directory of its own file, and /root/reference is read-only.
    ``oracle/Makefile`` (target ``ref``) compiles pyastar/astar.cpp where it lies
    into ``oracle/_ref/pyastar/astar.so`` and puts a *symlink* to the reference's
    pyastar.py beside it, so the unmodified binding runs.
  * ``colour``: absent from the image; only feeds render grey levels
    (Simulation/utility.py:1,88-111).
  * ``np.random.choice`` / ``random.randint``: redirected to the Philox RESET
    stream for the duration of ``reset()`` so all three sides draw the same
    numbers in the same order (oracle/philox.py).
"""
from __future__ import annotations

import os
import subprocess
import sys
import types
from contextlib import contextmanager

import numpy as np

from . import philox

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BUILD = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("WF_REFERENCE_ROOT", "/root/reference")
if not os.path.isfile(os.path.join(REF_ROOT, "Simulation", "forest_fire.py")):
    # the GPU box: no reference tree, but `make -C oracle ref` (run by build() in the build container) left a snapshot of
    # the unmodified Simulation/ package in the git-ignored oracle/_ref/, which travels with the working tree
    REF_ROOT = REF_BUILD


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "Simulation", "forest_fire.py"))


def build_ref() -> str:
    """Compile the reference's A* (pyastar/astar.cpp) into oracle/_ref/pyastar/astar.so."""
    subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)
    return os.path.join(REF_BUILD, "pyastar", "astar.so")


def _install_shims():
    """Make ``from pyastar import pyastar`` and ``from colour import Color`` resolvable."""
    so = os.path.join(REF_BUILD, "pyastar", "astar.so")
    if not os.path.isfile(so) or not os.path.isfile(os.path.join(REF_BUILD, "pyastar", "pyastar.py")):
        build_ref()
    # oracle/_ref/pyastar/pyastar.py is the reference's own binding (copied there by `make ref`);
    # it looks for astar.so beside its own path, i.e. in oracle/_ref/pyastar.
    if REF_BUILD not in sys.path:
        sys.path.insert(0, REF_BUILD)
    if "colour" not in sys.modules:
        col = types.ModuleType("colour")
        rgb = {"green": (0.0, 128 / 255, 0.0), "red": (1.0, 0.0, 0.0), "black": (0.0, 0.0, 0.0),
               "brown": (165 / 255, 42 / 255, 42 / 255), "blue": (0.0, 0.0, 1.0)}

        class Color:
            def __init__(self, name):
                self.red, self.green, self.blue = rgb[name.lower()]

        col.Color = Color
        sys.modules["colour"] = col


def load_reference():
    """Import the reference's Simulation package (from REF_ROOT) and return its modules."""
    _install_shims()
    if REF_ROOT not in sys.path:
        sys.path.insert(sys.path.index(REF_BUILD) + 1, REF_ROOT)  # after oracle/_ref (the reference's pyastar has no astar.so)
    import Simulation.constants as constants  # noqa: E402
    import Simulation.utility as utility  # noqa: E402
    import Simulation.environment as environment  # noqa: E402
    import Simulation.forest_fire as forest_fire  # noqa: E402
    return constants, utility, environment, forest_fire


@contextmanager
def philox_rng(stream: philox.ResetStream):
    """Route the reference's random draws to the shared RESET stream."""
    import random as pyrandom
    old_choice, old_randint = np.random.choice, pyrandom.randint

    def choice(a, *args, **kwargs):
        assert not args and not kwargs
        if isinstance(a, (int, np.integer)):
            a = range(int(a))
        return stream.choice(list(a))

    np.random.choice = choice
    pyrandom.randint = stream.randint
    try:
        yield
    finally:
        np.random.choice = old_choice
        pyrandom.randint = old_randint


_WALK_FN = None


def reference_walk_policy():
    """``DQN.choose_randomwalk_action`` (DQN.py:353-389) as a plain function of ``self``.

    DQN.py imports Keras at module level, which the image lacks, so the method's source is taken
    from the reference file with ``ast`` and compiled on its own; nothing is copied into the repo."""
    global _WALK_FN
    if _WALK_FN is None:
        import ast
        src = open(os.path.join(REF_ROOT, "DQN.py")).read()
        tree = ast.parse(src)
        fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "choose_randomwalk_action")
        mod = ast.Module(body=[fn], type_ignores=[])
        ns = {"np": np}
        exec(compile(mod, os.path.join(REF_ROOT, "DQN.py"), "exec"), ns)
        _WALK_FN = ns["choose_randomwalk_action"]
    return _WALK_FN


class _PolicyStream:
    """np.random.choice stand-in reading the POLICY stream of one (env, episode, step)."""

    def __init__(self, seed, env_id, episode, t):
        self.a = (seed, env_id, episode, t)
        self.j = 0

    def choice(self, seq):
        seq = list(seq)
        u = philox.policy_draw(*self.a, self.j)
        self.j += 1
        return seq[u % len(seq)]

    def randint(self, a, b):
        raise AssertionError("the walk policy only uses np.random.choice")


DEFAULT_META = dict(
    death_penalty=-1000, contained_bonus=1000, default_reward=-1,
    wind=[0.54, (0, 0)], n_actions=4, a_speed=1, make_rivers=False,
    containment_wins=False, allow_dig_toggle=False,
)


class RefEnv:
    """One reference ``ForestFire`` driven by the shared stream.

    ``cfg`` keys follow Simulation/constants.py:30-47 (``width``/``height``/``wind``/
    ``a_speed``/``make_rivers``/``allow_dig_toggle``/...), plus ``seed`` and
    ``extra_ignitions`` (ignitions applied through the public
    ``World.set_fire_to``, environment.py:233, right after ``reset()``).
    """

    def __init__(self, cfg: dict, env_id: int = 0):
        constants, utility, environment, forest_fire = load_reference()
        self.mod_env = environment
        self.meta = constants.METADATA
        self.cfg = dict(DEFAULT_META)
        self.cfg.update(cfg)
        self.env_id = env_id
        self.seed = int(self.cfg.get("seed", 0))
        self.episode = -1
        self.t = 0
        self._apply_meta()
        # World.__init__ runs one reset() (environment.py:183): feed it a throw-away stream
        with philox_rng(philox.ResetStream(self.seed ^ 0x5EED, env_id, 0xFFFFFFFF)):
            self.sim = forest_fire.ForestFire()
        self.sim.width, self.sim.height = self.W, self.H

    def _apply_meta(self):
        c = self.cfg
        self.W, self.H = int(c["width"]), int(c["height"])
        m = self.meta
        for k in ("death_penalty", "contained_bonus", "default_reward", "n_actions", "a_speed",
                  "make_rivers", "containment_wins", "allow_dig_toggle"):
            m[k] = c[k]
        m["width"], m["height"] = self.W, self.H
        m["wind"] = c["wind"] if c["wind"] == "random" else [c["wind"][0], tuple(c["wind"][1])]
        m["a_speed_iter"] = getattr(self, "a_iter", c["a_speed"])
        # WIDTH/HEIGHT are frozen at import (environment.py:22-23); they are plain
        # module globals, so re-pointing them re-sizes the next World().
        self.mod_env.WIDTH, self.mod_env.HEIGHT = self.W, self.H

    def reset(self):
        self._apply_meta()
        self.episode += 1
        self.t = 0
        with philox_rng(philox.ResetStream(self.seed, self.env_id, self.episode)):
            obs = self.sim.reset()
        for k in range(int(self.cfg.get("extra_ignitions", 0))):
            u0, u1 = philox.ignite_draw(self.seed, self.env_id, self.episode, k)
            self.sim.W.set_fire_to((u0 % self.W, u1 % self.H))
            obs = self.sim.W.get_state()
        return obs

    def random_action(self) -> int:
        return philox.action_draw(self.seed, self.env_id, self.episode, self.t) % int(self.cfg["n_actions"])

    def walk_action(self) -> int:
        """The reference's own heuristic policy code, fed from the shared POLICY stream."""
        self._apply_meta()
        fake_self = types.SimpleNamespace(sim=self.sim)
        with philox_rng(_PolicyStream(self.seed, self.env_id, self.episode, self.t)):
            return int(reference_walk_policy()(fake_self))

    def step(self, action):
        self._apply_meta()  # METADATA is a process-wide global in the reference
        out = self.sim.step(action)
        self.a_iter = self.meta["a_speed_iter"]  # Q8: lives in METADATA, survives reset()
        self.t += 1
        return out

    # ---- canonical state dump (same planes the C-ABI's wf_get_state returns) ----
    def planes(self):
        from Simulation.utility import layer
        env = self.sim.W.env
        W = self.sim.W
        typ = env[:, :, layer["type"]].astype(np.uint8)
        burning = np.zeros((self.W, self.H), np.uint8)
        for (x, y) in W.burning_cells:
            burning[x, y] = 1
        fm_inf = np.isinf(env[:, :, layer["fire_mobility"]]).astype(np.uint8)
        fuel = env[:, :, layer["fuel"]].astype(np.int32)
        temp = env[:, :, layer["temp"]].astype(np.float64)
        apos = env[:, :, layer["agent_pos"]].astype(np.uint8)
        alive = 1 if W.agents else 0
        ax, ay = (int(W.agents[0].x), int(W.agents[0].y)) if W.agents else (-1, -1)
        return dict(type=typ, burning=burning, fm_inf=fm_inf, fuel=fuel, temp=temp, apos=apos,
                    alive=alive, ax=ax, ay=ay, fire_at_border=int(bool(W.fire_at_border)),
                    running=int(bool(W.RUNNING)), wind_speed=float(W.wind_speed),
                    wind_x=int(W.wind_vector[0]), wind_y=int(W.wind_vector[1]))
