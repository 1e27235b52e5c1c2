// hsfm_warp.cu -- instantiations and launch dispatch of the warp-per-world symmetric-pair kernel (hsfm_warp.cuh).
#include "hsfm_warp.cuh"

namespace pms {
namespace {

template <typename F> cudaError_t dispatch(bool split, int T, int minb, F &&f)
{
#define PMS_WARP_CASE(TT)                                                                                   \
    case TT:                                                                                                \
        if (split) return minb == 2 ? f(hsfm_warp_kernel<double, true, TT, 2>, WarpSmem<double, true>::bytes(32 * TT))   \
                                    : f(hsfm_warp_kernel<double, true, TT, 3>, WarpSmem<double, true>::bytes(32 * TT));  \
        return minb == 2 ? f(hsfm_warp_kernel<float, false, TT, 2>, WarpSmem<float, false>::bytes(32 * TT))              \
                         : f(hsfm_warp_kernel<float, false, TT, 3>, WarpSmem<float, false>::bytes(32 * TT));
    switch (T) {
        PMS_WARP_CASE(1)
        PMS_WARP_CASE(2)
        PMS_WARP_CASE(3)
        PMS_WARP_CASE(4)
    default: return cudaErrorInvalidValue;
    }
#undef PMS_WARP_CASE
}

} // namespace

cudaError_t warp_kernel_prepare(bool split, int T, int minb, int wpc, int *blocks_per_sm, size_t *smem_bytes)
{
    return dispatch(split, T, minb, [&](auto kern, size_t per_world) {
        const size_t smem = per_world * wpc;
        *smem_bytes = smem;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kern, (wpc + 1) * 32, smem);
    });
}

cudaError_t warp_kernel_launch(bool split, int T, int minb, const HsfmArgs &a, int blocks, int threads, size_t smem,
                               cudaStream_t stream)
{
    return dispatch(split, T, minb, [&](auto kern, size_t) {
        kern<<<blocks, threads, smem, stream>>>(a);
        return cudaGetLastError();
    });
}

} // namespace pms
