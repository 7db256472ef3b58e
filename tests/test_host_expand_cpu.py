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


@pytest.mark.parametrize("switch", ["WF_HOST_NO_AVX512", "WF_HOST_NO_AVX2"])
def test_expand_code_paths_agree(switch):
    """The expansion picks AVX-512 / AVX2 / PDEP / table once per process: run the other paths in a child process (same
    inputs, odd and even grid sizes, so both the cache-line-aligned and the unaligned store loops are taken)."""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np, sys\n"
        "sys.path.insert(0, %r)\n"
        "from wildfire_control_python_b200 import _lib\n"
        "L = _lib.lib()\n"
        "for n, W, H in ((33, 14, 14), (5, 13, 11), (7, 32, 32), (64, 10, 10)):\n"
        "    rng = np.random.default_rng(n)\n"
        "    epw = 2 if W <= 16 else 1\n"
        "    bits = W * H * 3\n"
        "    records, rw = (n + epw - 1) // epw, (epw * bits + 31) // 32\n"
        "    obs = (rng.random((n, bits)) < 0.4).astype(np.uint8)\n"
        "    packed = np.zeros((records, rw), np.uint32)\n"
        "    for r in range(records):\n"
        "        b = obs[r * epw:(r + 1) * epw].reshape(-1)\n"
        "        p = np.zeros(rw * 32, np.uint8); p[:len(b)] = b\n"
        "        packed[r] = np.packbits(p, bitorder='little').view('<u4')\n"
        "    out = np.full((n, bits), 7, np.uint8)\n"
        "    assert L.wf_expand_packed_obs(packed.ctypes.data, out.ctypes.data, n, W, H, 2) == 0\n"
        "    assert np.array_equal(out, obs), (n, W, H)\n"
        "print('ok')\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **{switch: "1"}), capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
