// hsfm_duo_api.cuh -- host-side entry points of the pair-warp / owner-warp kernel (instantiated in hsfm_duo.cu).
#pragma once
#include "pms_common.cuh"

namespace pms {

constexpr int DUO_MAX_TILES = 2;            // N <= 64

// one world per CTA of 96 T + 32 threads (T = ceil(N/32) tiles: 2T pair warps + T owner warps + 1 service warp); comp = compensated hi + lo state.
// prepare: shared-memory size and resident CTAs per SM; launch: grid = n_worlds
// sens = the instantiation that serves a sensor plan (a.sp)
cudaError_t duo_kernel_prepare(bool comp, bool sens, int T, int *blocks_per_sm, size_t *smem_bytes);
cudaError_t duo_kernel_launch(bool comp, bool sens, int T, const HsfmArgs &a, int n_worlds, size_t smem, cudaStream_t stream);

} // namespace pms
