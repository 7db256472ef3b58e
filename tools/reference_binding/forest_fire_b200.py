import ctypes as C
import numpy as np
from os.path import abspath, dirname, join
from .constants import METADATA
from .utility import grass

lib = C.cdll.LoadLibrary(join(dirname(abspath(__file__)), "libwildfire_b200.so"))

class WfConfig(C.Structure):                       # include/wildfire.h: wf_config
    _fields_ = [(n, C.c_int32) for n in ("width", "height", "n_actions", "a_speed", "allow_dig_toggle",
                "make_rivers", "containment_wins", "wind_random", "wind_x", "wind_y", "fuel", "radius",
                "extra_ignitions", "auto_reset")] + \
               [(n, C.c_double) for n in ("wind_speed", "death_penalty", "contained_bonus", "default_reward",
                "heat", "threshold")] + [("seed", C.c_uint64), ("env_id_base", C.c_int64)]

i32 = np.ctypeslib.ndpointer(dtype=np.int32, ndim=1, flags="C_CONTIGUOUS")
u8 = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
f64 = np.ctypeslib.ndpointer(dtype=np.float64, ndim=1, flags="C_CONTIGUOUS")
lib.wf_create.argtypes = [C.POINTER(WfConfig), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
lib.wf_destroy.argtypes = [C.c_void_p]
lib.wf_reset_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, u8, C.c_int32]
lib.wf_step_host.argtypes = [C.c_void_p, i32, u8, C.c_int32, f64, u8]
lib.wf_last_error.restype = C.c_char_p

class ForestFire:                                   # same surface as Simulation/forest_fire.py:18-54, batched
    def __init__(self, n_envs=1, seed=0, device=0):
        cfg = WfConfig()
        lib.wf_default_config(C.byref(cfg), METADATA["width"])
        cfg.height = METADATA["height"]; cfg.n_actions = METADATA["n_actions"]; cfg.a_speed = METADATA["a_speed"]
        cfg.allow_dig_toggle = METADATA["allow_dig_toggle"]; cfg.make_rivers = METADATA["make_rivers"]
        if METADATA["wind"] == "random": cfg.wind_random = 1
        else: cfg.wind_speed, (cfg.wind_x, cfg.wind_y) = METADATA["wind"][0], METADATA["wind"][1]
        cfg.death_penalty, cfg.contained_bonus = METADATA["death_penalty"], METADATA["contained_bonus"]
        cfg.default_reward = METADATA["default_reward"]
        cfg.heat, cfg.fuel, cfg.threshold = grass["heat"], grass["fuel"], grass["threshold"]
        cfg.seed = seed
        self.h = C.c_void_p(); self.n = n_envs
        if lib.wf_create(C.byref(cfg), n_envs, device, C.byref(self.h)): raise RuntimeError(lib.wf_last_error())
        W, H = cfg.width, cfg.height
        self.obs = np.zeros((n_envs, W, H, 3), np.uint8)          # World.get_state layout, x slow
        self.reward = np.zeros(n_envs); self.done = np.zeros(n_envs, np.uint8)

    def __del__(self):
        if getattr(self, "h", None): lib.wf_destroy(self.h); self.h = None

    def reset(self):                                # forest_fire.py:52-54
        if lib.wf_reset_host(self.h, None, None, self.obs, 0): raise RuntimeError(lib.wf_last_error())
        return self.obs.astype(np.float64)

    def step(self, actions):                        # forest_fire.py:30-49, one action per env
        a = np.ascontiguousarray(actions, np.int32)
        if lib.wf_step_host(self.h, a, self.obs, 0, self.reward, self.done): raise RuntimeError(lib.wf_last_error())
        return [self.obs.astype(np.float64), self.reward, self.done.astype(bool), {}]
