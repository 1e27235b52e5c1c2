// hsfm_duo_api.cuh -- host-side entry points of the pair-warp / owner-warp kernel (instantiated in hsfm_duo.cu).
#pragma once
#include "pms_common.cuh"

namespace pms {

constexpr int DUO_MAX_TILES = 2;            // N <= 64
// pair warps of a world: the 2T units of the pair scheme, `upw` units per pair warp.  upw = 2 evaluates two units side by side
// in one warp (two independent chains; partner loads and loop shared): fewer instructions and fewer warps per world -- one
// tile: always (the pair warp still finishes before the owner warp: 1024 x 32 1.04 -> 1.00 ms); two tiles: when it lets the
// batch fit ONE wave of CTAs it would not fit otherwise (5 instead of 4 worlds per SM; 512 x 64 1.48 -> 1.41 ms, 740 x 64
// 2.21 -> 1.77), else one unit per warp (shorter chains: 128 x 64 0.79 vs 0.91 ms, 1024 x 64 2.52 vs 2.69).
__host__ __device__ constexpr int duo_pair_warps(int T, int upw) { return 2 * T / upw; }
__host__ __device__ constexpr int duo_threads(int T, int upw) { return 32 * (duo_pair_warps(T, upw) + T + 1); }   // + T owner warps + 1 service warp
__host__ __device__ constexpr int duo_min_blocks(int T, int upw) { return T == 1 ? (upw == 2 ? 9 : 7) : (upw == 2 ? 5 : 4); }

// one world per CTA of duo_threads(T, upw) threads (T = ceil(N/32) tiles: pair warps + T owner warps + 1 service warp); comp = compensated hi + lo state.
// prepare: shared-memory size and resident CTAs per SM; launch: grid = n_worlds
// sens = the instantiation that serves a sensor plan (a.sp)
cudaError_t duo_kernel_prepare(bool comp, bool sens, int T, int upw, int *blocks_per_sm, size_t *smem_bytes);
cudaError_t duo_kernel_launch(bool comp, bool sens, int T, int upw, const HsfmArgs &a, int n_worlds, size_t smem, cudaStream_t stream);

} // namespace pms
