"""ctypes binding of the C oracle (oracle/wf_oracle.c) -- TEST INFRASTRUCTURE.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libwf_oracle.so")


class Config(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("n_actions", C.c_int32), ("a_speed", C.c_int32),
        ("allow_dig_toggle", C.c_int32), ("make_rivers", C.c_int32), ("containment_wins", C.c_int32),
        ("wind_random", C.c_int32), ("wind_x", C.c_int32), ("wind_y", C.c_int32), ("fuel", C.c_int32),
        ("radius", C.c_int32), ("extra_ignitions", C.c_int32), ("_pad", C.c_int32),
        ("wind_speed", C.c_double), ("death_penalty", C.c_double), ("contained_bonus", C.c_double),
        ("default_reward", C.c_double), ("heat", C.c_double), ("threshold", C.c_double),
        ("seed", C.c_uint64),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "wf_oracle.c")
    if force or not os.path.isfile(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", HERE, "libwf_oracle.so"], check=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        p = C.c_void_p
        u8p, i32p, f64p = (C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_double))
        L.wfo_default_config.argtypes = [C.POINTER(Config), C.c_int]
        L.wfo_create.restype = p
        L.wfo_create.argtypes = [C.POINTER(Config), C.c_int64]
        L.wfo_destroy.argtypes = [p]
        L.wfo_reset.argtypes = [p]
        L.wfo_reset_at.argtypes = [p, C.c_int, C.c_int]
        L.wfo_step.restype = C.c_int
        L.wfo_step.argtypes = [p, C.c_int, u8p, f64p, C.POINTER(C.c_int)]
        L.wfo_stream_action.restype = C.c_int
        L.wfo_stream_action.argtypes = [p]
        L.wfo_policy_action.restype = C.c_int
        L.wfo_policy_action.argtypes = [p]
        L.wfo_get_obs.argtypes = [p, u8p]
        L.wfo_get_planes.argtypes = [p, u8p, u8p, u8p, i32p, f64p, u8p]
        L.wfo_get_scalars.argtypes = [p, i32p]
        L.wfo_get_wind_speed.restype = C.c_double
        L.wfo_get_wind_speed.argtypes = [p]
        L.wfo_get_coef.argtypes = [p, f64p]
        L.wfo_set_a_speed_iter.argtypes = [p, C.c_int]
        L.wfo_set_fire_to.argtypes = [p, C.c_int, C.c_int]
        L.wfo_set_planes.argtypes = [p, u8p, u8p, u8p, i32p, f64p]
        L.wfo_set_agent.argtypes = [p] + [C.c_int] * 6
        L.wfo_rollout.restype = C.c_int64
        L.wfo_rollout.argtypes = [C.POINTER(Config), C.c_int64, C.c_int, C.c_int, C.c_int, f64p,
                                  C.POINTER(C.c_int64)]
        L.wfo_batch_create.restype = p
        L.wfo_batch_create.argtypes = [C.POINTER(Config), C.c_int64, C.c_int, C.c_int]
        L.wfo_batch_destroy.argtypes = [p]
        L.wfo_batch_step.restype = C.c_int64
        L.wfo_batch_step.argtypes = [p, C.c_int, f64p, C.POINTER(C.c_int64)]
        L.wfo_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        _lib = L
    return _lib


def make_config(cfg: dict) -> Config:
    """``cfg`` uses the reference's METADATA key names (constants.py:30-47)."""
    c = Config()
    size = int(cfg.get("width", cfg.get("size", 10)))
    lib().wfo_default_config(C.byref(c), size)
    c.width = int(cfg.get("width", size))
    c.height = int(cfg.get("height", size))
    wind = cfg.get("wind", [0.54, (0, 0)])
    if wind == "random":
        c.wind_random = 1
    else:
        c.wind_speed = float(wind[0])
        c.wind_x, c.wind_y = int(wind[1][0]), int(wind[1][1])
    for k in ("n_actions", "a_speed", "allow_dig_toggle", "make_rivers", "containment_wins", "fuel",
              "radius", "extra_ignitions"):
        if k in cfg:
            setattr(c, k, int(cfg[k]))
    for k in ("death_penalty", "contained_bonus", "default_reward", "heat", "threshold"):
        if k in cfg:
            setattr(c, k, float(cfg[k]))
    c.seed = int(cfg.get("seed", 0))
    return c


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


class OracleEnv:
    """One oracle environment; same call surface as ``oracle.ref_harness.RefEnv``."""

    def __init__(self, cfg: dict, env_id: int = 0):
        self.cfg = make_config(cfg)
        self.W, self.H = self.cfg.width, self.cfg.height
        self.h = lib().wfo_create(C.byref(self.cfg), env_id)

    def __del__(self):
        if getattr(self, "h", None):
            lib().wfo_destroy(self.h)
            self.h = None

    def reset(self, start=None):
        if start is None:
            lib().wfo_reset(self.h)
        else:
            lib().wfo_reset_at(self.h, int(start[0]), int(start[1]))
        return self.obs()

    def obs(self):
        o = np.empty((self.W, self.H, 3), np.uint8)
        lib().wfo_get_obs(self.h, _ptr(o, C.c_uint8))
        return o

    def random_action(self) -> int:
        return lib().wfo_stream_action(self.h)

    def walk_action(self) -> int:
        """DQN.choose_randomwalk_action (DQN.py:353-389) with POLICY-stream draws."""
        return lib().wfo_policy_action(self.h)

    def step(self, action: int):
        o = np.empty((self.W, self.H, 3), np.uint8)
        r = C.c_double()
        d = C.c_int()
        rc = lib().wfo_step(self.h, int(action), _ptr(o, C.c_uint8), C.byref(r), C.byref(d))
        if rc != 0:
            raise IndexError("list index out of range (step on an env without agent)")
        return o, r.value, bool(d.value), {}

    def set_a_speed_iter(self, v):
        lib().wfo_set_a_speed_iter(self.h, int(v))

    def set_fire_to(self, x, y):
        lib().wfo_set_fire_to(self.h, int(x), int(y))

    def scalars(self):
        s = np.zeros(16, np.int32)
        lib().wfo_get_scalars(self.h, _ptr(s, C.c_int32))
        return s

    def coef(self):
        c = np.zeros(4, np.float64)
        lib().wfo_get_coef(self.h, _ptr(c, C.c_double))
        return c

    def planes(self):
        W, H = self.W, self.H
        typ = np.empty((W, H), np.uint8)
        burning = np.empty((W, H), np.uint8)
        fm_inf = np.empty((W, H), np.uint8)
        fuel = np.empty((W, H), np.int32)
        temp = np.empty((W, H), np.float64)
        apos = np.empty((W, H), np.uint8)
        lib().wfo_get_planes(self.h, _ptr(typ, C.c_uint8), _ptr(burning, C.c_uint8), _ptr(fm_inf, C.c_uint8),
                             _ptr(fuel, C.c_int32), _ptr(temp, C.c_double), _ptr(apos, C.c_uint8))
        s = self.scalars()
        return dict(type=typ, burning=burning, fm_inf=fm_inf, fuel=fuel, temp=temp, apos=apos,
                    alive=int(s[0]), ax=int(s[1]), ay=int(s[2]), fire_at_border=int(s[6]),
                    running=int(s[5]), wind_speed=lib().wfo_get_wind_speed(self.h),
                    wind_x=int(s[11]), wind_y=int(s[12]))

    def set_planes(self, type=None, burning=None, fm_inf=None, fuel=None, temp=None):
        def c(a, dt):
            return None if a is None else np.ascontiguousarray(a, dtype=dt)
        t, b, f, fu, te = c(type, np.uint8), c(burning, np.uint8), c(fm_inf, np.uint8), c(fuel, np.int32), c(temp, np.float64)
        lib().wfo_set_planes(self.h, _ptr(t, C.c_uint8), _ptr(b, C.c_uint8), _ptr(f, C.c_uint8),
                             _ptr(fu, C.c_int32), _ptr(te, C.c_double))

    def set_agent(self, alive, ax, ay, visible=1, dead=0, digging=1):
        lib().wfo_set_agent(self.h, int(alive), int(ax), int(ay), int(visible), int(dead), int(digging))


def rollout(cfg: dict, n_envs: int, n_steps: int, n_threads: int = 0, env_id_base: int = 0):
    """Throughput driver: returns (env_steps, episodes, checksum)."""
    c = make_config(cfg)
    chk = C.c_double()
    eps = C.c_int64()
    n = lib().wfo_rollout(C.byref(c), env_id_base, n_envs, n_steps, n_threads, C.byref(chk), C.byref(eps))
    return int(n), int(eps.value), float(chk.value)


class OracleBatch:
    """Persistent batch of oracle envs stepped with stream actions by ``n_threads`` host threads."""

    def __init__(self, cfg: dict, n_envs: int, n_threads: int = 1, env_id_base: int = 0):
        self.cfg = make_config(cfg)
        self.n_envs, self.n_threads = n_envs, max(1, min(n_threads, n_envs))
        self.h = lib().wfo_batch_create(C.byref(self.cfg), env_id_base, n_envs, self.n_threads)
        self.checksum = C.c_double(0.0)
        self.episodes = C.c_int64(0)

    def step(self, n_steps: int = 1) -> int:
        return int(lib().wfo_batch_step(self.h, n_steps, C.byref(self.checksum), C.byref(self.episodes)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().wfo_batch_destroy(self.h)
            self.h = None


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*[int(x) & 0xFFFFFFFF for x in ctr])
    k = (C.c_uint32 * 2)(*[int(x) & 0xFFFFFFFF for x in key])
    o = (C.c_uint32 * 4)()
    lib().wfo_philox4x32_10(c, k, o)
    return list(o)
