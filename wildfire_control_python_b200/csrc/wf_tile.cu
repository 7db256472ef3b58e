// wf_tile.cu -- "tile" kernel family: grids wider or taller than 32 cells (256x256, 1024x1024).
#include "wf_families.cuh"

namespace wf {
struct TileState { int dummy; };
int tile_extra_planes() { return 0; }
cudaError_t tile_create(TileState** out, const DevState&, const StepCfg&) { *out = new TileState(); return cudaSuccess; }
void tile_destroy(TileState* t) { delete t; }
cudaError_t launch_tile_family(TileState*, const DevState&, const StepCfg&, const TileIO&, cudaStream_t, int64_t*) {
    return cudaErrorNotSupported;
}
cudaError_t tile_after_set_state(TileState*, const DevState&, const StepCfg&, cudaStream_t, int64_t*) { return cudaSuccess; }
}  // namespace wf
