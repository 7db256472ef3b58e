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


def _encode_change_blocks(prev, new, status, W, H, max_entries=13):
    """NumPy statement of the kernel's side of wf_host_session mode 2 (wf_common.cuh, wf_warp.cu): per record (e envs) the
    list of changed elements, or -- more than 13 of them -- the whole bit stream in the full area; four records per block."""
    n = prev.shape[0]
    epw = 2 if W <= 16 else 1
    bits = W * H * 3
    records = (n + epw - 1) // epw
    rec_words = (epw * bits + 31) // 32
    full_stride = (rec_words + 3) // 4 * 4
    blocks = np.zeros(((records + 3) // 4, 32), np.uint32)
    full = np.full((records, full_stride), 0xDEADBEEF, np.uint32)  # must not be read unless flagged
    p, q = prev.reshape(n, bits), new.reshape(n, bits)
    for b in range(blocks.shape[0]):
        entries, mask = [], 0
        for w in range(4):
            r = 4 * b + w
            if r >= records:
                break
            envs = range(r * epw, min(n, (r + 1) * epw))
            st = 0
            for k, e in enumerate(envs):
                st |= int(status[e]) << (16 * k)
            ch = [(k * bits + int(i), int(q[e, i])) for k, e in enumerate(envs) for i in np.flatnonzero(p[e] != q[e])]
            if len(ch) > max_entries:
                mask |= 1 << w
                st |= 0x8000
                stream = np.zeros(full_stride * 32, np.uint8)
                flat = q[list(envs)].reshape(-1)
                stream[:len(flat)] = flat
                full[r] = np.packbits(stream, bitorder="little").view("<u4")
            else:
                entries += [((w * epw * bits + i) << 1) | v for i, v in ch]
            blocks[b, 1 + w] = st
        assert len(entries) <= 52
        blocks[b, 0] = len(entries) | (mask << 8)
        e16 = blocks[b].view(np.uint16)
        e16[10:10 + len(entries)] = entries
        e16[10 + len(entries):] = 0xFFFF
    return blocks, full, full_stride


@pytest.mark.parametrize("n_envs,W,H,threads", [(4096, 14, 14, 4), (37, 14, 14, 3), (9, 10, 10, 1), (21, 32, 32, 5), (6, 17, 13, 2)])
def test_apply_change_blocks_patches_the_array_in_place(n_envs, W, H, threads):
    """Host half of the persistent-observation session (wf_hostpool.cpp): after applying one step's blocks the caller's array
    equals the new observation, rewards / done flags decode from the status words, unflagged full-area records are not read."""
    rng = np.random.default_rng(n_envs + W)
    prev = (rng.random((n_envs, W, H, 3)) < 0.4).astype(np.uint8)
    new = prev.copy()
    bits = W * H * 3
    for e in range(n_envs):  # most envs change in a few elements, some in many (a reset), some not at all
        k = int(rng.choice([0, 1, 3, 6, 9, 14, 40, bits]))
        idx = rng.choice(bits, size=k, replace=False)
        new.reshape(n_envs, bits)[e, idx] ^= 1
    kind = rng.integers(0, 5, n_envs)
    done = rng.integers(0, 2, n_envs)
    count = rng.integers(0, W * H + 1, n_envs)
    status = kind | (done << 3) | (count << 4)
    blocks, full, stride = _encode_change_blocks(prev, new, status, W, H)
    obs = prev.copy()
    rew = np.zeros(n_envs)
    dn = np.zeros(n_envs, np.uint8)
    L = _lib.lib()
    _lib.check(L.wf_apply_change_blocks(blocks.ctypes.data, full.ctypes.data, stride, obs.ctypes.data, rew.ctypes.data, dn.ctypes.data,
                                        n_envs, W, H, -1.0, -1000.0, 1000.0, threads))
    assert np.array_equal(obs, new)
    assert np.array_equal(dn, done.astype(np.uint8))
    want = np.select([kind == 1, kind == 2, kind == 3, kind == 4], [-1.0, -1000.0, 1000.0, 1000.0 * (count / (W * H))], 0.0)
    assert np.array_equal(rew, want)
