// hsfm_duo_api.cuh -- host-side entry points of the pair-warp / owner-warp kernel (instantiated in hsfm_duo.cu).
#pragma once
#include "pms_common.cuh"

namespace pms {

constexpr int DUO_MAX_TILES = 2;            // N <= 64
// pair warps of a world: the 2T units of the pair scheme, `upw` units per pair warp.  upw = 2 evaluates two units side by side
// in one warp (two independent chains; partner loads and loop shared): fewer instructions and fewer warps per world.  One
// tile: always.  Two tiles -- `variant` of the entry points below: 1 = one unit per warp (7 warps, 4 worlds per SM, balanced
// roles); 2 = two units per warp (5 warps, 5 worlds per SM); 3 / 4 / 5 = two units per warp compiled for 7 / 4 / 3 worlds per SM
// (56 / 96 / 127 registers).  Which one is fastest depends on how the batch fills the SMs (pms_api.cu: duo_model_ms).
__host__ __device__ constexpr int duo_pair_warps(int T, int upw) { return 2 * T / upw; }
__host__ __device__ constexpr int duo_threads(int T, int upw) { return 32 * (duo_pair_warps(T, upw) + T + 1); }   // + T owner warps + 1 service warp
__host__ __device__ constexpr int duo_min_blocks(int T, int upw) { return T == 1 ? (upw == 2 ? 10 : 7) : (upw == 2 ? 5 : 4); }   // one tile: 64 registers (measured 9 / 10 / 12 / 14 worlds per SM: 1.007 / 0.988 / 1.048 / 1.169 ms on 1024 x 32)
constexpr int DUO_VARIANTS = 5;
__host__ __device__ constexpr int duo_variant_upw(int variant) { return variant >= 2 ? 2 : 1; }
__host__ __device__ constexpr int duo_variant_blocks(int variant) { return variant == 1 ? 4 : (variant == 2 ? 5 : (variant == 3 ? 7 : (variant == 4 ? 4 : 3))); }

// one world per CTA of duo_threads(T, upw) threads (T = ceil(N/32) tiles: pair warps + T owner warps + 1 service warp); comp = compensated hi + lo state.
// prepare: shared-memory size and resident CTAs per SM; launch: grid = n_worlds
// sens = the instantiation that serves a sensor plan (a.sp)
// variant: 1 .. 5 as above (one tile: always two units per warp, whatever is passed); a sensor plan runs variants 3 .. 5 as 2
cudaError_t duo_kernel_prepare(bool comp, bool sens, int T, int variant, int *blocks_per_sm, size_t *smem_bytes);
cudaError_t duo_kernel_launch(bool comp, bool sens, int T, int variant, const HsfmArgs &a, int n_worlds, size_t smem, cudaStream_t stream);

} // namespace pms
