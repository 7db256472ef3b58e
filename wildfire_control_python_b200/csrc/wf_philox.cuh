// wf_philox.cuh -- Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy
// as 1, 2, 3", SC'11), the counter-based generator behind the shared random stream.
// Product-side implementation; the oracle has its own (oracle/wf_oracle.c, oracle/philox.py) and
// both are pinned to the Random123 known-answer vectors (tests/test_philox.py, tests/test_parity_gpu.py).
//
// Stream contract (identical on all sides):
//   key = (seed & 0xffffffff, seed >> 32);  counter = (env_id, episode, index, stream)
//   stream 0 RESET : sequential draws of one World.reset(); draw k = word (k & 3) of block k >> 2
//   stream 1 ACTION: step t of the episode = word (t & 3) of block t >> 2; action = word % n_actions
//   stream 2 IGNITE: index = k-th extra ignition; cell = (word0 % W, word1 % H)
//   stream 3 POLICY: heuristic walk policy, step t: draw j (< 12) = word (j & 3) of block 3t + (j >> 2)
//   stream 4 EXPLORE: eps-greedy of the in-kernel Q-network: step t = words 2(t&1), 2(t&1)+1 of block t >> 1
#pragma once
#include <stdint.h>

namespace wf {

constexpr uint32_t kStreamReset = 0u, kStreamAction = 1u, kStreamIgnite = 2u, kStreamPolicy = 3u, kStreamExplore = 4u;

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

}  // namespace wf
