"""Shared helpers of the `-m gpu` parity tests: compare a BatchedForestFire with oracle envs."""
import numpy as np
import torch

from oracle import wf_oracle as wo
from wildfire_control_python_b200 import _lib
from wildfire_control_python_b200.batched import BatchedForestFire

TEMP_TOL = 1e-9  # |temp_gpu - temp_oracle| on grass cells; counters are exact, so only float64 summation order differs


def make_pair(n_envs, cfg, auto_reset=False, obs_dtype=torch.uint8):
    """A GPU batch and the matching list of oracle envs (same seed, env ids 0..n-1)."""
    gpu = BatchedForestFire(n_envs, obs_dtype=obs_dtype, auto_reset=auto_reset, **cfg)
    orc = [wo.OracleEnv(cfg, env_id=i) for i in range(n_envs)]
    return gpu, orc


def to_np(t):
    return t.detach().cpu().numpy()


def compare_states(tag, gpu, orc, envs=None, obs=None):
    st = {k: to_np(v) for k, v in gpu.get_state().items()}
    sc = st["scalars"]
    obs_np = None if obs is None else to_np(obs)
    for i in (range(len(orc)) if envs is None else envs):
        p = orc[i].planes()
        t = f"{tag} env {i}"
        for k in ("type", "burning", "fm_inf", "apos"):
            assert np.array_equal(st[k][i], p[k]), f"{t}: plane {k}\ngpu=\n{st[k][i].T}\noracle=\n{p[k].T}"
        assert np.array_equal(st["fuel"][i].astype(np.int32), np.maximum(p["fuel"], 0)), f"{t}: fuel"
        assert sc[i, _lib.S_ALIVE] == p["alive"], f"{t}: alive"
        if p["alive"]:
            assert (sc[i, _lib.S_AX], sc[i, _lib.S_AY]) == (p["ax"], p["ay"]), f"{t}: agent xy"
        assert sc[i, _lib.S_FIRE_AT_BORDER] == p["fire_at_border"], f"{t}: fire_at_border"
        assert sc[i, _lib.S_RUNNING] == p["running"], f"{t}: running"
        assert (sc[i, _lib.S_WIND_X], sc[i, _lib.S_WIND_Y]) == (p["wind_x"], p["wind_y"]), f"{t}: wind vector"
        assert gpu.wind_speed_table[sc[i, _lib.S_WIND_ID]] == p["wind_speed"], f"{t}: wind speed"
        assert sc[i, _lib.S_N_BURNING] == int(p["burning"].sum()), f"{t}: n_burning"
        grass = p["type"] == 0
        err = np.abs(st["temp"][i] - p["temp"])[grass].max(initial=0.0)
        assert err <= TEMP_TOL, f"{t}: temp err {err}"
        if obs_np is not None:
            want = orc[i].obs()
            assert np.array_equal(obs_np[i].astype(np.uint8), want), f"{t}: obs"
            assert set(np.unique(obs_np[i])) <= {0, 1}
