"""Host half of wf_step_host's packed-observation path (wf_hostpool.cpp), testable without a GPU:
``wf_expand_packed_obs`` must be the inverse of packing the [N][W][H][3] 0/1 array into per-group bit streams."""
import ctypes as C

import numpy as np
import pytest

from wildfire_control_python_b200 import _lib


@pytest.mark.parametrize("n_envs,W,H,threads", [(16, 14, 14, 1), (4097, 14, 14, 4), (5, 10, 10, 3), (7, 32, 32, 2),
                                                  (3, 17, 13, 8), (64, 16, 16, 16)])
def test_expand_packed_obs_inverts_bit_packing(n_envs, W, H, threads):
    rng = np.random.default_rng(n_envs * 7 + W)
    obs = (rng.random((n_envs, W, H, 3)) < 0.3).astype(np.uint8)
    epw = 2 if W <= 16 else 1
    env_bits = W * H * 3
    records = (n_envs + epw - 1) // epw
    rec_words = (epw * env_bits + 31) // 32
    packed = np.zeros((records, rec_words), np.uint32)
    flat = obs.reshape(n_envs, env_bits)
    for r in range(records):
        bits = flat[r * epw:(r + 1) * epw].reshape(-1)
        padded = np.zeros(rec_words * 32, np.uint8)
        padded[:len(bits)] = bits
        packed[r] = np.packbits(padded, bitorder="little").view("<u4")
    out = np.full((n_envs, W, H, 3), 9, np.uint8)
    L = _lib.lib()
    _lib.check(L.wf_expand_packed_obs(packed.ctypes.data, out.ctypes.data, n_envs, W, H, threads))
    assert np.array_equal(out, obs)


def test_expand_rejects_bad_arguments():
    L = _lib.lib()
    buf = np.zeros(64, np.uint32)
    out = np.zeros(64, np.uint8)
    assert L.wf_expand_packed_obs(buf.ctypes.data, out.ctypes.data, 1, 40, 40, 1) == _lib.WF_ERR_INVALID
    assert L.wf_expand_packed_obs(None, out.ctypes.data, 1, 10, 10, 1) == _lib.WF_ERR_INVALID
