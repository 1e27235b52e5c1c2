// hsfm_warp_api.cuh -- host-side entry points of the warp-per-world kernel family (instantiated in hsfm_warp.cu).
#pragma once
#include "pms_common.cuh"

namespace pms {

constexpr int WARP_MAX_TILES = 4;           // N <= 128
constexpr int WARP_MAX_WPC = 8;             // world warps per CTA (256 threads)

// shared-memory opt-in + occupancy of the instantiation (split = MIXED precision, T = ceil(N/32) tiles,
// cfg = register budget, see WarpCfg in hsfm_warp.cuh; cfg 1 and 2 allow at most 4 worlds per CTA) for `wpc` worlds per CTA;
// comp = FP32 mode with compensated (hi + lo) state; sens = the instantiation that serves a sensor plan (a.sp; register
// budget 0 only)
cudaError_t warp_kernel_prepare(bool split, bool comp, bool sens, int T, int cfg, int wpc, int *blocks_per_sm, size_t *smem_bytes);
cudaError_t warp_kernel_launch(bool split, bool comp, bool sens, int T, int cfg, const HsfmArgs &a, int blocks, int threads, size_t smem,
                               cudaStream_t stream);

} // namespace pms
