// hsfm_warp.cu -- instantiations and launch dispatch of the warp-per-world symmetric-pair kernel (hsfm_warp.cuh).
#include "hsfm_warp.cuh"

namespace pms {
namespace {

// comp: FP32 mode with the state carried as compensated hi + lo sums (PMS_FP32); plain fp32 state is PMS_FP32_PLAIN
// sens: the launch serves a sensor plan (pms_sensor_schedule) -- a separate instantiation, register budget 0 only (the
// plan's code is compiled out of the default kernels: carrying it switched off cost 11 % on C4)
template <typename F> cudaError_t dispatch(bool split, bool comp, bool sens, int T, int cfg, F &&f)
{
#define PMS_WARP_INST(TT, CF)                                                                               \
    (split ? f(hsfm_warp_kernel<double, true, TT, CF>, WarpSmem<double, true>::bytes(32 * TT))              \
     : comp ? f(hsfm_warp_kernel<float, false, TT, CF, true>, WarpSmem<float, false, true>::bytes(32 * TT)) \
            : f(hsfm_warp_kernel<float, false, TT, CF>, WarpSmem<float, false>::bytes(32 * TT)))
#define PMS_WARP_SENS(TT)                                                                                           \
    (split ? f(hsfm_warp_kernel<double, true, TT, 0, false, true>, WarpSmem<double, true>::bytes(32 * TT))          \
     : comp ? f(hsfm_warp_kernel<float, false, TT, 0, true, true>, WarpSmem<float, false, true>::bytes(32 * TT))    \
            : f(hsfm_warp_kernel<float, false, TT, 0, false, true>, WarpSmem<float, false>::bytes(32 * TT)))
#define PMS_WARP_CASE(TT)                                                                                   \
    case TT:                                                                                                \
        if (sens) return PMS_WARP_SENS(TT);                                                                 \
        switch (cfg) {                                                                                      \
        case 1: return PMS_WARP_INST(TT, 1);                                                                \
        case 2: return PMS_WARP_INST(TT, 2);                                                                \
        case 3: return PMS_WARP_INST(TT, 3);                                                                \
        default: return PMS_WARP_INST(TT, 0);                                                               \
        }
    switch (T) {
        PMS_WARP_CASE(1)
        PMS_WARP_CASE(2)
        PMS_WARP_CASE(3)
        PMS_WARP_CASE(4)
    default: return cudaErrorInvalidValue;
    }
#undef PMS_WARP_CASE
#undef PMS_WARP_SENS
#undef PMS_WARP_INST
}

} // namespace

cudaError_t warp_kernel_prepare(bool split, bool comp, bool sens, int T, int cfg, int wpc, int *blocks_per_sm, size_t *smem_bytes)
{
    return dispatch(split, comp, sens, T, cfg, [&](auto kern, size_t per_world) {
        const size_t smem = per_world * wpc;
        *smem_bytes = smem;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kern, wpc * 32, smem);
    });
}

cudaError_t warp_kernel_launch(bool split, bool comp, bool sens, int T, int cfg, const HsfmArgs &a, int blocks, int threads, size_t smem,
                               cudaStream_t stream)
{
    return dispatch(split, comp, sens, T, cfg, [&](auto kern, size_t) {
        kern<<<blocks, threads, smem, stream>>>(a);
        return cudaGetLastError();
    });
}

} // namespace pms
