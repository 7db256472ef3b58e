// wf_families.cuh -- launch interfaces of the two kernel families (warp: W,H <= 32; tile: larger).
#pragma once
#include "wf_common.cuh"

namespace wf {

// Q-network of WF_POLICY_MLP, device pointers (owned by the handle).
struct MlpPolicy {
    const float* w1;    // [n_in][hid], Keras orientation
    const float* base;  // [hid]: bias1 + sum over cells of w1[(cell, channel 2)] -- the all-free, no-fire, no-agent map
    const float* w2;    // [hid][n_actions]
    const float* b2;    // [n_actions]
    int32_t hid, n_actions;
    uint32_t eps_u32;   // explore iff a 32-bit draw is below this
};

struct WarpIO {
    const int32_t* actions;  // [K][N] or nullptr (ACTION stream)
    void* obs;               // [K][N][W][H][3] or nullptr
    double* reward;          // [K][N] or nullptr
    uint8_t* done;           // [K][N] or nullptr
    const uint8_t* mask;     // reset mode: [N] or nullptr
    const wf_init* init;     // reset mode: [N] or nullptr
    int32_t obs_dtype, K, a_iter0, reset_mode;
    uint32_t magicH;         // ceil(2^32 / H)
    int32_t policy;          // actions == nullptr: WF_POLICY_STREAM, WF_POLICY_WALK or WF_POLICY_MLP
    int32_t* actions_out;    // [K][N] or nullptr: the actions the policy chose
    MlpPolicy mlp;           // WF_POLICY_MLP only
};
cudaError_t launch_warp_family(const DevState& s, const StepCfg& c, const WarpIO& io, cudaStream_t stream);

struct TileIO {
    const int32_t* actions;  // [K][N] or nullptr (ACTION stream / policy)
    void* obs;               // [K][N][W][H][3] or nullptr
    double* reward;          // [K][N] or nullptr
    uint8_t* done;           // [K][N] or nullptr
    const uint8_t* mask;     // reset mode: [N] or nullptr
    const wf_init* init;     // reset mode: [N] or nullptr
    int32_t obs_dtype, K, a_iter0, reset_mode;
    int32_t policy;          // actions == nullptr: WF_POLICY_STREAM or WF_POLICY_WALK
    int32_t* actions_out;    // [K][N] or nullptr
};
struct TileState;
int tile_extra_planes();
cudaError_t tile_create(TileState** out, const DevState& s, const StepCfg& c);
void tile_destroy(TileState* t);
void tile_geometry(const TileState* t, int32_t* threads, int32_t* cluster);
cudaError_t launch_tile_family(TileState* t, const DevState& s, const StepCfg& c, const TileIO& io,
                               cudaStream_t stream, int64_t* launches);
cudaError_t tile_after_set_state(TileState* t, const DevState& s, const StepCfg& c, cudaStream_t stream,
                                 int64_t* launches);

}  // namespace wf
