// wf_tile.cu -- "tile" kernel family: grids wider or taller than 32 cells (256x256, 1024x1024 ...).
//
// Same bit-plane state as the warp family (wf_common.cuh) but resident in HBM: a row x of H cells is
// HW = ceil(H/32) words, one thread owns one word (32 cells) per tick.  Per tick a word reads
//   * the heat-source mask S (1 bit/cell: burning and fuel >= 2) of itself and its 4 neighbours,
//   * its grass / fire / burning / fm_inf words,
// and ONLY IF it burns or receives heat the fuel planes and the hit counters of the heated cells.
// It writes the next source mask (ping-pong), whatever changed, and 96 bytes of observation.
// The per-burning-cell Python loop (forest_fire.py:85-106) is boolean algebra on words; the A*
// containment search (environment.py:342-377) is a persistent "reach" plane R (cells with a finite
// 4-connected path to a finite border point) that is re-flooded only when a dig may disconnect it.
//
// One step = agent_kernel (1 thread/env: move/dig/reap + local articulation test; envs whose reach
//                          plane may have been disconnected by the dig are appended to a work list)
//          -> flood_list_kernel (persistent CTAs drain the list: re-flood R; usually the list is empty)
//          -> tile_tick_kernel (the stencil: fuel, heat, ignition, per-env reductions)
//          -> finish_kernel (1 thread/env: reward/done/latch; finished envs go to the reset list)
//          -> reset_list_kernel (persistent CTAs drain the list: World.reset; usually empty)
//          -> obs_kernel  (World.get_state of every env: 2 bits/cell in, 3 bytes/cell out)
#include "wf_families.cuh"

namespace wf {

constexpr int kTileThreads = 256;

struct TileState {
    int32_t* acc;         // [N][4]: burning cells, grass cells, ignition-on-edge flag, burning-touches-reach flag
    int32_t* need_flood;  // [N] flag: R of this env must be re-flooded
    int32_t* flood_list;  // [N] env ids appended by agent_kernel, + counter
    int32_t* reset_list;  // [N] env ids appended by finish_kernel / mask_to_list_kernel, + counter
    int32_t* counters;    // [0] = flood_list length, [1] = reset_list length
    int32_t cur;          // which of the two S planes holds the sources of the NEXT tick
    int32_t P_S0, P_S1, P_R;
    int32_t flood_smem_ok;
};

int tile_extra_planes() { return 3; }

__device__ __forceinline__ uint32_t valid_word(int H, int w) {
    const int rem = H - 32 * w;
    return rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
}

// Literal border_points (environment.py:215-222): (x,0) (x,H-1) for all x; (0,y) (H-1,y) for all y.
__device__ __forceinline__ uint32_t seed_word(int W, int H, int x, int w) {
    const uint32_t v = valid_word(H, w);
    if (x == 0 || x == H - 1) return v;
    uint32_t m = 0u;
    if (w == 0) m |= 1u;
    if (w == (H - 1) >> 5) m |= 1u << ((H - 1) & 31);
    return m & v;
}
__device__ __forceinline__ uint32_t edge_word(int W, int H, int x, int w) {
    const uint32_t v = valid_word(H, w);
    if (x == 0 || x == W - 1) return v;
    uint32_t m = 0u;
    if (w == 0) m |= 1u;
    if (w == (H - 1) >> 5) m |= 1u << ((H - 1) & 31);
    return m & v;
}

__device__ __forceinline__ uint32_t& plane_word(const DevState& s, int p, int env, int x, int w) {
    return s.planes[word_index(s, p, env, x, w)];
}
__device__ __forceinline__ bool get_bit(const DevState& s, int p, int env, int x, int y) {
    return (s.planes[word_index(s, p, env, x, y >> 5)] >> (y & 31)) & 1u;
}

// Agent.dig (environment.py:123-133) on planes in HBM + incremental maintenance of the reach plane.
__device__ void tile_dig(const DevState& s, const TileState& t, int env, int x, int y, const int32_t* sc) {
    const int w = y >> 5;
    const uint32_t bit = 1u << (y & 31);
    uint32_t& D = plane_word(s, P_D, env, x, w);
    if (D & bit) return;
    plane_word(s, P_G, env, x, w) &= ~bit;
    plane_word(s, P_F, env, x, w) &= ~bit;
    plane_word(s, P_BT, env, x, w) &= ~bit;
    plane_word(s, P_WT, env, x, w) &= ~bit;
    D |= bit;
    uint32_t& I = plane_word(s, P_I, env, x, w);
    const bool was_free = !(I & bit);
    I |= bit;
    if (!was_free) return;
    // get_reward only searches for a path while `not fire_at_border and len(border_points)`
    // (environment.py:345): once either latch is set R is never read again in this episode
    // (World.reset re-floods it), so it is not maintained.
    if (sc[WF_S_FIRE_AT_BORDER] || sc[WF_S_LATCHED]) return;
    uint32_t& R = plane_word(s, t.P_R, env, x, w);
    if (!(R & bit)) return;  // the cell had no path to the border: nobody reached the border through it
    R &= ~bit;
    // Does removing this cell possibly disconnect its neighbours from the border?  Its free 4-neighbours
    // were all in R.  If they stay connected to each other through the ring of 8 surrounding cells,
    // every path through the dug cell can be re-routed and R is unchanged elsewhere ("simple point").
    const int W = s.W, H = s.H;
    const bool on_seed = (x == 0 || x == H - 1 || y == 0 || y == H - 1);
    bool f[8];  // N, NE, E, SE, S, SW, W, NW  (screen coordinates: N = y-1, E = x+1)
    const int dx[8] = {0, 1, 1, 1, 0, -1, -1, -1}, dy[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int nx = x + dx[k], ny = y + dy[k];
        f[k] = (nx >= 0 && nx < W && ny >= 0 && ny < H) && !get_bit(s, P_I, env, nx, ny);
    }
    int n4 = 0, groups = 0;
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
        if (!f[k]) continue;
        n4++;
        // this 4-neighbour starts a new group unless it is ring-connected to the previous 4-neighbour
        const int pk = (k + 6) & 7, pd = (k + 7) & 7;
        if (!(f[pk] && f[pd])) groups++;
    }
    if (n4 == 4 && groups == 0) groups = 1;  // full ring
    if ((on_seed ? n4 > 0 : groups > 1) && !t.need_flood[env]) {
        t.need_flood[env] = 1;
        t.flood_list[atomicAdd(&t.counters[0], 1)] = env;
    }
}

// World.set_fire_to (environment.py:233-246) on planes in HBM (keeps the source mask consistent).
__device__ void tile_set_fire(const DevState& s, const TileState& t, int env, int x, int y, int32_t* sc) {
    const int w = y >> 5;
    const uint32_t bit = 1u << (y & 31);
    plane_word(s, P_G, env, x, w) &= ~bit;
    plane_word(s, P_BT, env, x, w) &= ~bit;
    plane_word(s, P_D, env, x, w) &= ~bit;
    plane_word(s, P_WT, env, x, w) &= ~bit;
    plane_word(s, P_F, env, x, w) |= bit;
    plane_word(s, P_B, env, x, w) |= bit;
    uint32_t ge2 = 0u;
    for (int q = 1; q < s.FB; ++q) ge2 |= plane_word(s, P_FU0 + q, env, x, w);
    if (ge2 & bit) plane_word(s, t.cur ? t.P_S1 : t.P_S0, env, x, w) |= bit;
    if (x == 0 || x == s.W - 1 || y == 0 || y == s.H - 1) sc[WF_S_FIRE_AT_BORDER] = 1;
}

// ---------------------------------------------------------------------------------------------
// Reach plane: R = finite cells 4-connected to a finite border point (flood from the border over
// ~fm_inf).  Whole CTA; in-place monotone relaxation until a full sweep changes nothing.
__device__ void flood_block(const DevState& s, const TileState& t, int env) {
    const int W = s.W, H = s.H, HW = s.HW, nwords = W * HW;
    uint32_t* R = &plane_word(s, t.P_R, env, 0, 0);
    const uint32_t* I = &plane_word(s, P_I, env, 0, 0);
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) {
        const int x = i / HW, w = i - x * HW;
        R[i] = seed_word(W, H, x, w) & ~I[i];
    }
    __syncthreads();
    for (;;) {
        int changed = 0;
        for (int i = threadIdx.x; i < nwords; i += blockDim.x) {
            const int x = i / HW, w = i - x * HW;
            const uint32_t free_ = ~I[i] & valid_word(H, w);
            if (!free_) continue;
            const uint32_t old = R[i];
            uint32_t n = old;
            if (x > 0) n |= R[i - HW];
            if (x < W - 1) n |= R[i + HW];
            if (w > 0) n |= R[i - 1] >> 31;
            if (w < HW - 1) n |= R[i + 1] << 31;
            n = hfill(n & free_, free_);
            if (n != old) {
                R[i] = n;
                changed = 1;
            }
        }
        if (!__syncthreads_or(changed)) break;
    }
    if (threadIdx.x == 0) t.need_flood[env] = 0;
}

// ForestFire.step part 1: the action (Agent.move :141-155, toggle_digging :136-138) and, on tick
// steps, Agent.is_dead (:116-120).  One thread per env.
__global__ void agent_kernel(DevState s, StepCfg c, TileState t, const int32_t* actions, int do_tick, int policy,
                             int32_t* actions_out) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env == 0) t.counters[1] = 0;  // the reset list was drained by the previous step
    if (env >= s.N) return;
    int32_t* sc = s.scal + (size_t)env * WF_NSCALARS;
    int32_t* acc = t.acc + 4 * env;
    acc[0] = acc[1] = acc[2] = acc[3] = 0;
    sc[WF_S_RESERVED] = sc[WF_S_RUNNING];  // "act": was running at step start (finished envs are frozen)
    if (!sc[WF_S_RUNNING]) return;
    int action;
    if (actions != nullptr) {
        action = actions[env];
    } else if (policy == WF_POLICY_WALK) {  // DQN.choose_randomwalk_action, DQN.py:353-389
        action = 0;
        if (sc[WF_S_ALIVE]) {
            const int mx = s.W / 2, my = s.H / 2, px = sc[WF_S_AX], py = sc[WF_S_AY];
            int a0 = 0, a1 = 0;
            if (px >= mx && py > my) { a0 = 1; a1 = 3; }
            if (px > mx && py <= my) { a0 = 1; a1 = 2; }
            if (px <= mx && py < my) { a0 = 0; a1 = 2; }
            if (px < mx && py >= my) { a0 = 0; a1 = 3; }
            uint32_t pw[4];
            for (int j = 0, count = 0;; ++j) {
                if ((j & 3) == 0)
                    philox4x32_10((uint32_t)(c.env_id_base + env), (uint32_t)sc[WF_S_EPISODE],
                                  3u * (uint32_t)sc[WF_S_T] + (uint32_t)(j >> 2), kStreamPolicy, c.key0, c.key1, pw);
                action = (pw[j & 3] & 1u) ? a1 : a0;
                const int nx = px + (action == 2 ? 1 : action == 3 ? -1 : 0);
                const int ny = py + (action == 1 ? 1 : action == 0 ? -1 : 0);
                const bool fire_at_loc = nx >= 0 && nx < s.W && ny >= 0 && ny < s.H && get_bit(s, P_F, env, nx, ny);
                if (!fire_at_loc || count > 10) break;
                count++;
            }
        }
    } else {
        uint32_t w[4];
        const uint32_t tt = (uint32_t)sc[WF_S_T];
        philox4x32_10((uint32_t)(c.env_id_base + env), (uint32_t)sc[WF_S_EPISODE], tt >> 2, kStreamAction, c.key0, c.key1, w);
        action = (int)(w[tt & 3u] % (uint32_t)c.n_actions);
    }
    if (actions_out != nullptr) actions_out[env] = action;
    int ax = sc[WF_S_AX], ay = sc[WF_S_AY];
    if (!sc[WF_S_ALIVE]) return;
    if (action >= 0 && action < 4) {
        sc[WF_S_VISIBLE] = 0;  // Q1
        const int nx = ax + (action == 2 ? 1 : action == 3 ? -1 : 0);
        const int ny = ay + (action == 1 ? 1 : action == 0 ? -1 : 0);
        if (nx >= 0 && nx < s.W && ny >= 0 && ny < s.H && !get_bit(s, P_WT, env, nx, ny)) {
            ax = nx; ay = ny;
            sc[WF_S_AX] = ax; sc[WF_S_AY] = ay; sc[WF_S_VISIBLE] = 1;
            const bool onfire = get_bit(s, P_F, env, nx, ny);
            if (sc[WF_S_DIGGING] && !onfire) tile_dig(s, t, env, nx, ny, sc);
            if (onfire) sc[WF_S_DEAD] = 1;
        }
    }
    if (c.allow_dig_toggle && action == 4) {
        sc[WF_S_DIGGING] ^= 1;
        if (sc[WF_S_DIGGING]) tile_dig(s, t, env, ax, ay, sc);
    }
    if (do_tick && (sc[WF_S_DEAD] || get_bit(s, P_F, env, ax, ay))) {
        sc[WF_S_VISIBLE] = 0;
        sc[WF_S_ALIVE] = 0;
        atomicAdd(&s.stats[ST_DEATHS], 1ull);
    }
}

// Persistent CTAs drain the flood list (empty on most steps).
__global__ void __launch_bounds__(1024) flood_list_kernel(DevState s, TileState t) {
    const int n = t.counters[0];
    for (int k = blockIdx.x; k < n; k += gridDim.x) {
        flood_block(s, t, t.flood_list[k]);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// The stencil.
// Active word (it burns or is heated): reduce_fuel :297-307, burn-out, apply_heat_from_to :278-294,
// set_fire_to :233-246.  Returns the word's heat sources for the next tick.  Planes this thread does
// not hold in registers are patched with fire-and-forget reductions (RED), so nothing waits on them.
template <int FB>
__device__ __forceinline__ uint32_t tick_active_word(uint32_t* P, size_t pstride, uint32_t& G, uint32_t& B, uint32_t h0,
                                                     uint32_t h1, uint32_t h2, uint32_t h3, const DevState& s,
                                                     const StepCfg& c, uint32_t* hrow, int wid, int kmin,
                                                     uint32_t edge, int& my_edge) {
    uint32_t m = h0 | h1 | h2 | h3;
    // issue the loads of this word together: fuel planes (only if it burns) and the first heated cells
    uint32_t FU[FB];
    if (B) {
#pragma unroll
        for (int q = 0; q < FB; ++q) FU[q] = P[(P_FU0 + q) * pstride];
    }
    int y[4];
    uint32_t v[4];
    uint32_t mm = m;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        y[j] = mm ? __ffs(mm) - 1 : -1;
        mm &= mm - 1u;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = y[j] >= 0 ? hrow[y[j]] : 0u;

    uint32_t out = 0u, ge2 = 0u;
    const bool burning_word = B != 0u;
    if (B) {
        uint32_t borrow = B;
#pragma unroll
        for (int q = 0; q < FB; ++q) {
            const uint32_t f = FU[q];
            FU[q] = f ^ borrow;
            borrow &= ~f;
        }
        uint32_t nz = 0u;
#pragma unroll
        for (int q = 0; q < FB; ++q) {
            FU[q] &= ~borrow;
            nz |= FU[q];
        }
        out = B & ~nz;
#pragma unroll
        for (int q = 0; q < FB; ++q) P[(P_FU0 + q) * pstride] = FU[q];
#pragma unroll
        for (int q = 1; q < FB; ++q) ge2 |= FU[q];  // fuel >= 2, for every cell of the word
    }
    uint32_t ign = 0u;
    auto heat_cell = [&](int yy, uint32_t old) {
        const uint32_t nv = old + (((h0 >> yy) & 1u) | (((h1 >> yy) & 1u) << 8) | (((h2 >> yy) & 1u) << 16) |
                                   (((h3 >> yy) & 1u) << 24));
        hrow[yy] = nv;
        const bool ig = kmin >= 0 ? (int)__dp4a(nv, 0x01010101u, 0u) >= kmin : ignites(nv, s.wind, wid, c.threshold);
        if (ig) ign |= 1u << yy;
    };
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (y[j] >= 0) heat_cell(y[j], v[j]);
    while (mm) {  // more than 4 heated cells in one word: rare
        const int yy = __ffs(mm) - 1;
        mm &= mm - 1u;
        heat_cell(yy, hrow[yy]);
    }
    if (out) {  // type := burnt whatever it was (a dug-while-burning cell is dirt, Q7)
        atomicOr(&P[P_BT * pstride], out);
        atomicAnd(&P[P_D * pstride], ~out);
        atomicAnd(&P[P_WT * pstride], ~out);
        atomicAnd(&P[P_F * pstride], ~out);
        G &= ~out; B &= ~out;
    }
    if (ign) {
        atomicOr(&P[P_F * pstride], ign);
        G &= ~ign; B |= ign;
        if (!burning_word) {  // heated-only word that ignites (rare): now its fuel planes are needed
#pragma unroll
            for (int q = 1; q < FB; ++q) ge2 |= P[(P_FU0 + q) * pstride];
        }
    }
    if (ign | out) {
        P[P_G * pstride] = G;
        P[P_B * pstride] = B;
    }
    if (ign & edge) my_edge = 1;
    return B & ge2;  // sources of the next tick: burning with fuel >= 2
}

// does a burning cell of this word sit in, or next to, the border-connected region R?
__device__ __forceinline__ bool touches_reach(const uint32_t* R, uint32_t B, int x, int w, int W, int HW) {
    uint32_t near = R[0];
    near |= (near << 1) | (near >> 1);
    if (x > 0) near |= R[-HW];
    if (x < W - 1) near |= R[HW];
    if (w > 0) near |= R[-1] >> 31;
    if (w < HW - 1) near |= R[1] << 31;
    return (B & near) != 0u;
}

// VW = words per thread along y in the streaming phase: 4 (128-bit loads; needs HW % 4 == 0) or 1.
// Phase 1 streams G, B and the source mask of every word and queues the ACTIVE words (burning or
// heated) of the CTA in shared memory; phase 2 hands one queued word to each thread, so the
// dependent loads of the active path (fuel planes, hit counters) run at full occupancy instead of
// serially inside the few threads that happen to own a fire front.
template <int FB, int VW>
__global__ void __launch_bounds__(kTileThreads) tile_tick_kernel(DevState s, StepCfg c, TileState t, int do_tick,
                                                                 int hw_shift) {
    constexpr int QCAP = kTileThreads * VW;
    __shared__ int red[4];
    __shared__ int q_n;
    __shared__ uint32_t q_idx[QCAP], q_G[QCAP], q_B[QCAP], q_h[4][QCAP];
    const int env = blockIdx.y;
    if (threadIdx.x < 4) red[threadIdx.x] = 0;
    if (threadIdx.x == 0) q_n = 0;
    __syncthreads();
    const int W = s.W, H = s.H, HW = s.HW, nwords = W * HW;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) * VW;  // first word of this thread
    const int32_t* sc = s.scal + (size_t)env * WF_NSCALARS;
    const bool act = sc[WF_S_RESERVED] != 0;
    const bool want_touch = !sc[WF_S_FIRE_AT_BORDER] && !sc[WF_S_LATCHED];
    const size_t pstride = (size_t)s.N * s.RS * s.HW;
    uint32_t* const P0 = s.planes + word_index(s, 0, env, 0, 0);
    const int scur = t.cur ? t.P_S1 : t.P_S0, snxt = t.cur ? t.P_S0 : t.P_S1;
    int my_nb = 0, my_ng = 0, my_edge = 0, my_touch = 0;
    if (i < nwords) {
        const int x = hw_shift >= 0 ? (i >> hw_shift) : (i / HW);
        const int w0 = i - x * HW;
        uint32_t* P = P0 + i;
        uint32_t G[VW], B[VW];
        if (VW == 4) {
            const uint4 g4 = *reinterpret_cast<const uint4*>(P + P_G * pstride);
            const uint4 b4 = *reinterpret_cast<const uint4*>(P + P_B * pstride);
            G[0] = g4.x; G[VW > 1 ? 1 : 0] = g4.y; G[VW > 2 ? 2 : 0] = g4.z; G[VW > 3 ? 3 : 0] = g4.w;
            B[0] = b4.x; B[VW > 1 ? 1 : 0] = b4.y; B[VW > 2 ? 2 : 0] = b4.z; B[VW > 3 ? 3 : 0] = b4.w;
        } else {
            G[0] = P[P_G * pstride];
            B[0] = P[P_B * pstride];
        }
        if (act && do_tick) {
            const uint32_t* Sc = P + (size_t)scur * pstride;
            uint32_t* Sn = P + (size_t)snxt * pstride;
            uint32_t S[VW + 2], Sup[VW], Sdn[VW];
            if (VW == 4) {
                const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                const uint4 s4 = *reinterpret_cast<const uint4*>(Sc);
                const uint4 u4 = x > 0 ? *reinterpret_cast<const uint4*>(Sc - HW) : z;
                const uint4 d4 = x < W - 1 ? *reinterpret_cast<const uint4*>(Sc + HW) : z;
                S[1] = s4.x; S[VW > 1 ? 2 : 1] = s4.y; S[VW > 2 ? 3 : 1] = s4.z; S[VW > 3 ? 4 : 1] = s4.w;
                Sup[0] = u4.x; Sup[VW > 1 ? 1 : 0] = u4.y; Sup[VW > 2 ? 2 : 0] = u4.z; Sup[VW > 3 ? 3 : 0] = u4.w;
                Sdn[0] = d4.x; Sdn[VW > 1 ? 1 : 0] = d4.y; Sdn[VW > 2 ? 2 : 0] = d4.z; Sdn[VW > 3 ? 3 : 0] = d4.w;
            } else {
                S[1] = Sc[0];
                Sup[0] = x > 0 ? Sc[-HW] : 0u;
                Sdn[0] = x < W - 1 ? Sc[HW] : 0u;
            }
            S[0] = w0 > 0 ? Sc[-1] : 0u;
            S[VW + 1] = w0 + VW < HW ? Sc[VW] : 0u;
#pragma unroll
            for (int k = 0; k < VW; ++k) {
                const uint32_t h0 = G[k] & ((S[k + 1] >> 1) | (S[k + 2] << 31));  // d = N (0,-1): source at y+1
                const uint32_t h1 = G[k] & ((S[k + 1] << 1) | (S[k] >> 31));      // d = S (0,+1): source at y-1
                const uint32_t h2 = G[k] & Sup[k];                                // d = E (+1,0): source at x-1
                const uint32_t h3 = G[k] & Sdn[k];                                // d = W (-1,0): source at x+1
                if (B[k] | h0 | h1 | h2 | h3) {
                    const int pos = atomicAdd(&q_n, 1);
                    q_idx[pos] = (uint32_t)(i + k);
                    q_G[pos] = G[k]; q_B[pos] = B[k];
                    q_h[0][pos] = h0; q_h[1][pos] = h1; q_h[2][pos] = h2; q_h[3][pos] = h3;
                } else {
                    my_ng += __popc(G[k]);  // inactive words cannot burn: B == 0
                }
            }
            // inactive words have no sources next tick; queued words overwrite their slot in phase 2
            if (VW == 4) *reinterpret_cast<uint4*>(Sn) = make_uint4(0u, 0u, 0u, 0u);
            else Sn[0] = 0u;
        } else {
#pragma unroll
            for (int k = 0; k < VW; ++k) {
                my_nb += __popc(B[k]);
                my_ng += __popc(G[k]);
                if (B[k] && want_touch && touches_reach(P + k + (size_t)t.P_R * pstride, B[k], x, w0 + k, W, HW)) my_touch = 1;
            }
        }
    }
    __syncthreads();
    // ---- phase 2: one queued active word per thread
    const int nq = q_n;
    if (nq) {
        const int wid = sc[WF_S_WIND_ID];
        const int kmin = s.wind->uniform[wid] ? s.wind->kmin[wid] : -1;
        for (int it = threadIdx.x; it < nq; it += blockDim.x) {
            const int wi = (int)q_idx[it];
            const int x = hw_shift >= 0 ? (wi >> hw_shift) : (wi / HW);
            const int w = wi - x * HW;
            uint32_t G = q_G[it], B = q_B[it];
            uint32_t* P = P0 + wi;
            const uint32_t sn = tick_active_word<FB>(P, pstride, G, B, q_h[0][it], q_h[1][it], q_h[2][it], q_h[3][it], s, c,
                                                     s.hits + ((size_t)env * W + x) * H + 32 * w, wid, kmin,
                                                     edge_word(W, H, x, w), my_edge);
            P[(size_t)snxt * pstride] = sn;
            my_nb += __popc(B);
            my_ng += __popc(G);
            if (B && want_touch && touches_reach(P + (size_t)t.P_R * pstride, B, x, w, W, HW)) my_touch = 1;
        }
    }
    // ---- per-env reductions: warp redux -> shared -> one atomic per block and quantity
    const unsigned FULL = 0xffffffffu;
    my_nb = __reduce_add_sync(FULL, my_nb);
    my_ng = __reduce_add_sync(FULL, my_ng);
    my_edge = __any_sync(FULL, my_edge);
    my_touch = __any_sync(FULL, my_touch);
    if ((threadIdx.x & 31) == 0) {
        if (my_nb) atomicAdd(&red[0], my_nb);
        if (my_ng) atomicAdd(&red[1], my_ng);
        if (my_edge) atomicOr(&red[2], 1);
        if (my_touch) atomicOr(&red[3], 1);
    }
    __syncthreads();
    if (threadIdx.x < 4 && red[threadIdx.x]) {
        int32_t* acc = t.acc + 4 * env;
        if (threadIdx.x < 2) atomicAdd(&acc[threadIdx.x], red[threadIdx.x]);
        else atomicOr(&acc[threadIdx.x], 1);
    }
}

// ---------------------------------------------------------------------------------------------
// World.get_state :399-402 for every env: [agent_pos, type == fire, fire_mobility != inf].
// One thread per word: 2 plane words in, 96 bytes (32 cells x 3 channels) out.
__global__ void __launch_bounds__(kTileThreads) obs_kernel(DevState s, void* obs, int obs_dtype) {
    __shared__ uint32_t spread3[256];
    for (int v = threadIdx.x; v < 256; v += blockDim.x) {
        uint32_t o = 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) o |= ((v >> i) & 1u) << (3 * i);
        spread3[v] = o;
    }
    __syncthreads();
    const int env = blockIdx.y;
    const int W = s.W, H = s.H, HW = s.HW, nwords = W * HW;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int ic = min(i, nwords - 1);  // tail threads stay alive for the warp shuffles
    const int x = ic / HW, w = ic - x * HW;
    const int32_t* sc = s.scal + (size_t)env * WF_NSCALARS;
    const size_t pstride = (size_t)s.N * s.RS * s.HW;
    const uint32_t* P = s.planes + word_index(s, 0, env, 0, 0) + ic;
    const uint32_t F = P[P_F * pstride];
    const uint32_t freerow = ~P[P_I * pstride] & valid_word(H, w);
    const uint32_t arow = (sc[WF_S_VISIBLE] && sc[WF_S_AX] == x && (sc[WF_S_AY] >> 5) == w) ? 1u << (sc[WF_S_AY] & 31) : 0u;
    const int ncell = min(32, H - 32 * w);
    const size_t e0 = (((size_t)env * W + x) * H + 32 * w) * 3;  // first output element of this word
    uint32_t r[3];
    {
        uint32_t p[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            p[k] = spread3[(arow >> (8 * k)) & 255u] | (spread3[(F >> (8 * k)) & 255u] << 1) |
                   (spread3[(freerow >> (8 * k)) & 255u] << 2);
        r[0] = p[0] | (p[1] << 24);
        r[1] = (p[1] >> 8) | (p[2] << 16);
        r[2] = (p[2] >> 16) | (p[3] << 8);
    }
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    // Fast path (warp-uniform): the warp's 32 words are 32 full words of ONE row-major run, so their
    // 32 x 96 output bytes are contiguous.  Chunk c (16 bytes) of that run comes from word c/6: fetch its
    // stream bits by shuffle so that every store instruction writes 512 contiguous bytes.
    const int i0 = i - lane;
    const bool fast = obs_dtype == WF_OBS_U8 && (H & 31) == 0 && i0 + 31 < nwords;
    if (fast) {
        uint8_t* base = static_cast<uint8_t*>(obs) + (((size_t)env * W) * H + (size_t)32 * i0) * 3;
        uint4* o = reinterpret_cast<uint4*>(base);
#pragma unroll
        for (int it = 0; it < 6; ++it) {
            const int cidx = it * 32 + lane;       // chunk index within the warp's 3072 bytes
            const int src = cidx / 6, k = cidx - 6 * src;
            const uint32_t a0 = __shfl_sync(FULL, r[0], src), a1 = __shfl_sync(FULL, r[1], src),
                           a2 = __shfl_sync(FULL, r[2], src);
            const uint32_t word = (k >> 1) == 0 ? a0 : (k >> 1) == 1 ? a1 : a2;
            const uint32_t bits = (word >> ((k & 1) * 16)) & 0xffffu;
            uint4 v;
            v.x = ((bits & 15u) * 0x00204081u) & 0x01010101u;
            v.y = (((bits >> 4) & 15u) * 0x00204081u) & 0x01010101u;
            v.z = (((bits >> 8) & 15u) * 0x00204081u) & 0x01010101u;
            v.w = (((bits >> 12) & 15u) * 0x00204081u) & 0x01010101u;
            o[cidx] = v;
        }
    } else if (i < nwords) {
        if (obs_dtype == WF_OBS_U8) {
            uint8_t* o8 = static_cast<uint8_t*>(obs) + e0;
            for (int b = 0; b < 3 * ncell; ++b) o8[b] = (uint8_t)((r[b >> 5] >> (b & 31)) & 1u);
        } else {
            float* of = static_cast<float*>(obs) + e0;
            for (int b = 0; b < 3 * ncell; ++b) of[b] = ((r[b >> 5] >> (b & 31)) & 1u) ? 1.0f : 0.0f;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// World.reset part 1 (reset_map, environment.py:59-67): every plane and the temp layer back to their
// defaults.  Grid = (slices, persistent CTAs over the reset list): a big env is initialised by several CTAs.
template <int FB>
__global__ void __launch_bounds__(1024) reset_init_kernel(DevState s, StepCfg c, TileState t) {
    const int n = t.counters[1];
    const int W = s.W, H = s.H, HW = s.HW, nwords = W * HW;
    const size_t pstride = (size_t)s.N * s.RS * s.HW;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    for (int k = blockIdx.y; k < n; k += gridDim.y) {
        const int env = t.reset_list[k];
        uint32_t* P0 = s.planes + word_index(s, 0, env, 0, 0);
        for (int i = tid; i < nwords; i += nthr) {
            const int x = i / HW, w = i - x * HW;
            const uint32_t valid = valid_word(H, w);
            uint32_t* P = P0 + i;
            P[P_G * pstride] = valid;
            P[P_F * pstride] = 0u; P[P_BT * pstride] = 0u; P[P_D * pstride] = 0u; P[P_WT * pstride] = 0u;
            P[P_B * pstride] = 0u; P[P_I * pstride] = 0u;
#pragma unroll
            for (int q = 0; q < FB; ++q) P[(P_FU0 + q) * pstride] = ((c.fuel >> q) & 1) ? valid : 0u;
            P[(size_t)t.P_S0 * pstride] = 0u;
            P[(size_t)t.P_S1 * pstride] = 0u;
            P[(size_t)t.P_R * pstride] = valid;  // open field: every cell reaches the border (re-flooded if rivers)
        }
        uint32_t* hits = s.hits + (size_t)env * W * H;  // temp layer := 0
        if (((size_t)env * W * H) % 4 == 0 && (W * H) % 4 == 0) {
            uint4* h4 = reinterpret_cast<uint4*>(hits);
            for (int i = tid; i < (W * H) / 4; i += nthr) h4[i] = make_uint4(0u, 0u, 0u, 0u);
        } else {
            for (int i = tid; i < W * H; i += nthr) hits[i] = 0u;
        }
    }
}

// World.reset part 2 (environment.py:186-212) of one env by one CTA: wind, river, fire, agent, ignitions.
template <int FB>
__device__ void reset_block(const DevState& s, const StepCfg& c, const TileState& t, const wf_init* init, int env) {
    __shared__ int sh_fab, sh_nb;
    const int W = s.W, H = s.H, HW = s.HW, nwords = W * HW;
    if (threadIdx.x == 0) { sh_fab = 0; sh_nb = 0; }
    __syncthreads();
    int32_t* sc = s.scal + (size_t)env * WF_NSCALARS;
    const uint32_t episode = (uint32_t)sc[WF_S_EPISODE] + 1u;  // every thread reads the old value ...
    __syncthreads();                                            // ... before thread 0 overwrites it
    if (threadIdx.x == 0) {
        ResetDraws dr((uint32_t)(c.env_id_base + env), episode, c.key0, c.key1);
        int wid = 0;
        if (c.wind_random) {  // :188-190
            const int si = dr.next() % 3u, wx = dr.next() % 3u, wy = dr.next() % 3u;
            wid = si * 9 + wx * 3 + wy;
        }
        const int cx = W / 2, cy = H / 2;
        if (c.make_rivers) {  // reset_map :69-95
            int river_x = dr.next() % (uint32_t)W;
            int river_y = 1 + dr.next() % 3u;
            while (river_y < H - (1 + (int)(dr.next() % 3u))) {
                const uint32_t bit = 1u << (river_y & 31);
                plane_word(s, P_G, env, river_x, river_y >> 5) &= ~bit;
                plane_word(s, P_WT, env, river_x, river_y >> 5) |= bit;
                plane_word(s, P_I, env, river_x, river_y >> 5) |= bit;
                const int new_y = river_y + 1;
                int new_x = river_x + ((dr.next() % 2u) ? -1 : 1);
                for (;;) {
                    const int lo = 1 + dr.next() % 3u;
                    bool chain = false;
                    if (lo <= new_x) chain = new_x < W - (1 + (int)(dr.next() % 3u));
                    if (chain || (new_x == cx && new_y == cy)) break;
                    new_x = river_x + ((dr.next() % 2u) ? -1 : 1);
                }
                river_x = new_x;
                river_y = new_y;
            }
        }
        tile_set_fire(s, t, env, cx, cy, sc);  // :203
        int ax, ay;
        if (init != nullptr && init[env].ax >= 0) {
            ax = init[env].ax; ay = init[env].ay;
        } else {
            const int rad = dr.next() % 3u;
            const int idx = dr.next() % (uint32_t)kCircleLen[rad];
            ax = cx + kCircle[rad][idx][0];
            ay = cy + kCircle[rad][idx][1];
        }
        {  // Agent.__init__ digs its start cell (:112-113); R is re-flooded below
            const uint32_t bit = 1u << (ay & 31);
            const int w = ay >> 5;
            plane_word(s, P_G, env, ax, w) &= ~bit; plane_word(s, P_F, env, ax, w) &= ~bit;
            plane_word(s, P_BT, env, ax, w) &= ~bit; plane_word(s, P_WT, env, ax, w) &= ~bit;
            plane_word(s, P_D, env, ax, w) |= bit; plane_word(s, P_I, env, ax, w) |= bit;
            plane_word(s, t.P_R, env, ax, w) &= ~bit;
        }
        sc[WF_S_ALIVE] = 1; sc[WF_S_AX] = ax; sc[WF_S_AY] = ay; sc[WF_S_DEAD] = 0; sc[WF_S_DIGGING] = 1;
        sc[WF_S_VISIBLE] = 1; sc[WF_S_RUNNING] = 1; sc[WF_S_LATCHED] = 0;
        sc[WF_S_EPISODE] = (int32_t)episode; sc[WF_S_T] = 0; sc[WF_S_WIND_ID] = wid;
        sc[WF_S_WIND_X] = s.wind->wx[wid]; sc[WF_S_WIND_Y] = s.wind->wy[wid];
        sc[WF_S_FIRE_AT_BORDER] = 0;  // :212
    }
    __syncthreads();
    // extra ignitions: World.set_fire_to after reset() (IGNITE stream).  set_fire_to is idempotent and
    // commutes with itself, so the k-th ignitions run in parallel with atomics.
    const int scur = t.cur ? t.P_S1 : t.P_S0;
    for (int k = threadIdx.x; k < c.extra_ignitions; k += blockDim.x) {
        uint32_t wd[4];
        philox4x32_10((uint32_t)(c.env_id_base + env), episode, (uint32_t)k, kStreamIgnite, c.key0, c.key1, wd);
        const int x = (int)(wd[0] % (uint32_t)W), y = (int)(wd[1] % (uint32_t)H), w = y >> 5;
        const uint32_t bit = 1u << (y & 31);
        atomicAnd(&plane_word(s, P_G, env, x, w), ~bit);
        atomicAnd(&plane_word(s, P_BT, env, x, w), ~bit);
        atomicAnd(&plane_word(s, P_D, env, x, w), ~bit);
        atomicAnd(&plane_word(s, P_WT, env, x, w), ~bit);
        atomicOr(&plane_word(s, P_F, env, x, w), bit);
        atomicOr(&plane_word(s, P_B, env, x, w), bit);
        if (c.fuel >= 2) atomicOr(&plane_word(s, scur, env, x, w), bit);
        if (x == 0 || x == W - 1 || y == 0 || y == H - 1) sh_fab = 1;
    }
    __syncthreads();
    // Without rivers the only blocked cell is the agent's start cell, and one cell cannot cut a
    // >= 10x10 grid: R = every free cell (set by reset_init_kernel).  With rivers: flood.
    if (c.make_rivers) flood_block(s, t, env);
    else if (threadIdx.x == 0) t.need_flood[env] = 0;
    int n = 0;
    const uint32_t* B = &plane_word(s, P_B, env, 0, 0);
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) n += __popc(B[i]);
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&sh_nb, n);
    __syncthreads();
    if (threadIdx.x == 0) {
        sc[WF_S_N_BURNING] = sh_nb;
        if (sh_fab) sc[WF_S_FIRE_AT_BORDER] = 1;
    }
}

// ForestFire.step part 3: RUNNING flag (forest_fire.py:105-106), World.get_reward
// (environment.py:342-390).  One thread per env; finished envs are appended to the reset list.
__global__ void finish_kernel(DevState s, StepCfg c, TileState t, double* reward, uint8_t* done, int do_tick) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env == 0) t.counters[0] = 0;  // the flood list was drained before the tick
    if (env >= s.N) return;
    int32_t* sc = s.scal + (size_t)env * WF_NSCALARS;
    const int32_t* acc = t.acc + 4 * env;
    const bool act = sc[WF_S_RESERVED] != 0;
    double rew = 0.0;
    if (act) {
        const bool anyB = acc[0] > 0;
        sc[WF_S_N_BURNING] = acc[0];
        if (do_tick) {
            if (acc[2]) sc[WF_S_FIRE_AT_BORDER] = 1;
            if (!sc[WF_S_ALIVE] || !anyB) sc[WF_S_RUNNING] = 0;
        }
        const bool check = !sc[WF_S_FIRE_AT_BORDER] && !sc[WF_S_LATCHED] && anyB;
        if (check && !acc[3]) {
            sc[WF_S_LATCHED] = 1;  // bonus paid once (Q4), tested before the death test
            rew = c.contained_bonus;
            atomicAdd(&s.stats[ST_CONTAINED], 1ull);
        } else if (!sc[WF_S_ALIVE]) {
            rew = c.death_penalty;
        } else if (!anyB) {
            rew = __dmul_rn(c.contained_bonus, __ddiv_rn((double)acc[1], (double)(s.W * s.H)));
        } else {
            rew = c.default_reward;
        }
        sc[WF_S_T] += 1;
        atomicAdd(&s.stats[ST_STEPS], 1ull);
        if (!sc[WF_S_RUNNING]) {
            atomicAdd(&s.stats[ST_EPISODES], 1ull);
            if (sc[WF_S_ALIVE]) atomicAdd(&s.stats[ST_BURNOUTS], 1ull);
        }
    }
    const bool is_done = !sc[WF_S_RUNNING];
    if (reward) reward[env] = rew;
    if (done) done[env] = is_done ? 1 : 0;
    if (c.auto_reset && act && is_done) t.reset_list[atomicAdd(&t.counters[1], 1)] = env;
}

// wf_reset: the reset list is the mask.
__global__ void mask_to_list_kernel(TileState t, const uint8_t* mask, int n) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env < n && (mask == nullptr || mask[env])) t.reset_list[atomicAdd(&t.counters[1], 1)] = env;
}
__global__ void zero_counter_kernel(TileState t, int which) { t.counters[which] = 0; }

// Persistent CTAs drain the reset list: World.reset (environment.py:186-212), one CTA per env at a time.
template <int FB>
__global__ void __launch_bounds__(1024) reset_list_kernel(DevState s, StepCfg c, TileState t, const wf_init* init) {
    const int n = t.counters[1];
    for (int k = blockIdx.x; k < n; k += gridDim.x) {
        reset_block<FB>(s, c, t, init, t.reset_list[k]);
        __syncthreads();
    }
}

// S := B & (fuel >= 2), R re-flooded, n_burning recounted (after wf_set_state / wf_set_fire_to)
__global__ void rebuild_kernel(DevState s, TileState t) {
    __shared__ int sh_nb;
    const int env = blockIdx.x;
    const int nwords = s.W * s.HW;
    const size_t pstride = (size_t)s.N * s.RS * s.HW;
    uint32_t* P0 = s.planes + word_index(s, 0, env, 0, 0);
    if (threadIdx.x == 0) sh_nb = 0;
    __syncthreads();
    int n = 0;
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) {
        uint32_t* P = P0 + i;
        uint32_t ge2 = 0u;
        for (int q = 1; q < s.FB; ++q) ge2 |= P[(size_t)(P_FU0 + q) * pstride];
        const uint32_t B = P[P_B * pstride];
        P[(size_t)(t.cur ? t.P_S1 : t.P_S0) * pstride] = B & ge2;
        n += __popc(B);
    }
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&sh_nb, n);
    __syncthreads();
    if (threadIdx.x == 0) s.scal[(size_t)env * WF_NSCALARS + WF_S_N_BURNING] = sh_nb;
    flood_block(s, t, env);
}

// ---------------------------------------------------------------------------------------------
cudaError_t tile_create(TileState** out, const DevState& s, const StepCfg&) {
    TileState* t = new TileState();
    t->cur = 0;
    t->P_S0 = 7 + s.FB;
    t->P_S1 = 8 + s.FB;
    t->P_R = 9 + s.FB;
    t->flood_smem_ok = 0;
    cudaError_t e;
    if ((e = cudaMalloc(&t->acc, (size_t)s.N * 4 * sizeof(int32_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&t->need_flood, (size_t)s.N * sizeof(int32_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&t->flood_list, (size_t)s.N * sizeof(int32_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&t->reset_list, (size_t)s.N * sizeof(int32_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&t->counters, 2 * sizeof(int32_t))) != cudaSuccess) return e;
    cudaMemset(t->counters, 0, 2 * sizeof(int32_t));
    cudaMemset(t->acc, 0, (size_t)s.N * 4 * sizeof(int32_t));
    cudaMemset(t->need_flood, 0, (size_t)s.N * sizeof(int32_t));
    *out = t;
    return cudaSuccess;
}

void tile_destroy(TileState* t) {
    if (!t) return;
    cudaFree(t->acc); cudaFree(t->need_flood); cudaFree(t->flood_list); cudaFree(t->reset_list); cudaFree(t->counters);
    delete t;
}

static int cta_threads(const DevState& s) { return s.W * s.HW >= 8192 ? 1024 : 256; }
static int list_grid(const DevState& s) { return s.N < 296 ? s.N : 296; }  // 2 persistent CTAs per SM
static int init_slices(const DevState& s) {  // CTAs that share one env's plane initialisation
    const int per = (s.W * s.H + 65535) / 65536;  // ~64K cells (256 KB of hit counters) per CTA
    return per < 1 ? 1 : (per > 16 ? 16 : per);
}

template <int FB>
static cudaError_t run_family(TileState* t, const DevState& s, const StepCfg& c, const TileIO& io, cudaStream_t st,
                              int64_t* launches) {
    const int nwords = s.W * s.HW;
    const dim3 grid((nwords + kTileThreads - 1) / kTileThreads, s.N);
    const int eb = (s.N + 127) / 128;
    if (io.reset_mode) {
        zero_counter_kernel<<<1, 1, 0, st>>>(*t, 1);
        mask_to_list_kernel<<<eb, 128, 0, st>>>(*t, io.mask, s.N);
        reset_init_kernel<FB><<<dim3(init_slices(s), list_grid(s)), 1024, 0, st>>>(s, c, *t);
        reset_list_kernel<FB><<<list_grid(s), 1024, 0, st>>>(s, c, *t, io.init);
        *launches += 4;
    } else {
        agent_kernel<<<eb, 128, 0, st>>>(s, c, *t, io.actions, io.do_tick, io.policy, io.actions_out);
        flood_list_kernel<<<list_grid(s), 1024, 0, st>>>(s, *t);
        int hw_shift = -1;
        for (int k = 0; k < 16; ++k)
            if ((1 << k) == s.HW) hw_shift = k;
        if (s.HW % 4 == 0) {
            const dim3 g4((nwords / 4 + kTileThreads - 1) / kTileThreads, s.N);
            tile_tick_kernel<FB, 4><<<g4, kTileThreads, 0, st>>>(s, c, *t, io.do_tick, hw_shift);
        } else {
            tile_tick_kernel<FB, 1><<<grid, kTileThreads, 0, st>>>(s, c, *t, io.do_tick, hw_shift);
        }
        if (io.do_tick) t->cur ^= 1;
        finish_kernel<<<eb, 128, 0, st>>>(s, c, *t, io.reward, io.done, io.do_tick);
        *launches += 4;
        if (c.auto_reset) {
            reset_init_kernel<FB><<<dim3(init_slices(s), list_grid(s)), 1024, 0, st>>>(s, c, *t);
            reset_list_kernel<FB><<<list_grid(s), 1024, 0, st>>>(s, c, *t, nullptr);
            *launches += 2;
        }
    }
    if (io.obs) {
        obs_kernel<<<grid, kTileThreads, 0, st>>>(s, io.obs, io.obs_dtype);
        *launches += 1;
    }
    return cudaGetLastError();
}

cudaError_t launch_tile_family(TileState* t, const DevState& s, const StepCfg& c, const TileIO& io,
                               cudaStream_t stream, int64_t* launches) {
    if (s.FB == 5) return run_family<5>(t, s, c, io, stream, launches);
    return run_family<8>(t, s, c, io, stream, launches);
}

cudaError_t tile_after_set_state(TileState* t, const DevState& s, const StepCfg&, cudaStream_t stream,
                                 int64_t* launches) {
    rebuild_kernel<<<s.N, cta_threads(s), 0, stream>>>(s, *t);
    *launches += 1;
    return cudaGetLastError();
}

}  // namespace wf
