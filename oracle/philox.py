"""Philox4x32-10 counter-based RNG -- TEST INFRASTRUCTURE (oracle side).

This is the oracle's own statement of the shared random stream; the product has
an independent CUDA implementation in
``wildfire_control_python_b200/csrc/wf_philox.cuh``.  Both are pinned to the
Random123 known-answer vectors in ``tests/test_philox.py``.

The reference draws its randomness from ``np.random.choice`` / ``random.randint``
(reference: Simulation/environment.py:75-93,189-190; Simulation/utility.py:70,75).
To let the reference, the oracle and the GPU consume *identical* randomness we
replace those draws by this stream:

    key        = (seed & 0xffffffff, seed >> 32)
    counter    = (env_id, episode, index, stream)
    stream 0   = RESET   sequential draws of one World.reset():  draw k is word
                 (k & 3) of the block with index = k >> 2
    stream 1   = ACTION  step t of the episode uses word (t & 3) of the block with index = t >> 2
    stream 2   = IGNITE  index = k-th extra ignition, words 0/1 -> (x, y)
    stream 3   = POLICY  draws of the heuristic walk policy (DQN.py:353-389) at step t:
                 draw j (j < 12) is word (j & 3) of the block with index = 3 * t + (j >> 2)
    stream 4   = EXPLORE eps-greedy of the in-kernel Q-network at step t: words 2*(t & 1), 2*(t & 1) + 1
                 of the block with index = t >> 1  (explore iff word < eps * 2^32; action = word' % n_actions)
    a draw u picks ``seq[u % len(seq)]``;  randint(a, b) -> a + u % (b - a + 1)
"""
from __future__ import annotations

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF

STREAM_RESET = 0
STREAM_ACTION = 1
STREAM_IGNITE = 2
STREAM_POLICY = 3
STREAM_EXPLORE = 4


def philox4x32_10(ctr, key):
    """One Philox4x32-10 block.  ``ctr``: 4 uint32, ``key``: 2 uint32 -> 4 uint32."""
    c = [int(x) & MASK for x in ctr]
    k = [int(x) & MASK for x in key]
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & MASK, (p0 >> 32) ^ c[3] ^ k[1], p0 & MASK]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return c


def seed_key(seed: int):
    return (seed & MASK, (seed >> 32) & MASK)


def draw(seed: int, env_id: int, episode: int, stream: int, k: int) -> int:
    """k-th sequential 32-bit draw of (env, episode, stream)."""
    return philox4x32_10((env_id, episode, k >> 2, stream), seed_key(seed))[k & 3]


def action_draw(seed: int, env_id: int, episode: int, t: int) -> int:
    return philox4x32_10((env_id, episode, t >> 2, STREAM_ACTION), seed_key(seed))[t & 3]


def policy_draw(seed: int, env_id: int, episode: int, t: int, j: int) -> int:
    return philox4x32_10((env_id, episode, 3 * t + (j >> 2), STREAM_POLICY), seed_key(seed))[j & 3]


def explore_draw(seed: int, env_id: int, episode: int, t: int):
    """eps-greedy draws of the in-kernel Q-network at step t: (explore word, random-action word)."""
    w = philox4x32_10((env_id, episode, t >> 1, STREAM_EXPLORE), seed_key(seed))
    return (w[2], w[3]) if (t & 1) else (w[0], w[1])


def ignite_draw(seed: int, env_id: int, episode: int, k: int):
    w = philox4x32_10((env_id, episode, k, STREAM_IGNITE), seed_key(seed))
    return w[0], w[1]


class ResetStream:
    """Sequential reader of the RESET stream of one (env, episode)."""

    def __init__(self, seed: int, env_id: int, episode: int):
        self.seed, self.env_id, self.episode = seed, env_id, episode
        self.k = 0

    def next_u32(self) -> int:
        u = draw(self.seed, self.env_id, self.episode, STREAM_RESET, self.k)
        self.k += 1
        return u

    def choice(self, seq):
        seq = list(range(seq)) if isinstance(seq, int) else list(seq)
        return seq[self.next_u32() % len(seq)]

    def randint(self, a: int, b: int) -> int:
        return a + self.next_u32() % (b - a + 1)
