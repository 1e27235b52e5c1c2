// hsfm_duo.cu -- instantiations and launch dispatch of the pair-warp / owner-warp kernel (hsfm_duo.cuh).
#include "hsfm_duo.cuh"

namespace pms {
namespace {

template <typename F> cudaError_t dispatch(bool comp, bool sens, int T, F &&f)
{
#define PMS_DUO_INST(TT)                                                                                          \
    (sens ? (comp ? f(hsfm_duo_kernel<TT, true, true>, DuoSmem<TT>::BYTES) : f(hsfm_duo_kernel<TT, false, true>, DuoSmem<TT>::BYTES)) \
          : (comp ? f(hsfm_duo_kernel<TT, true, false>, DuoSmem<TT>::BYTES) : f(hsfm_duo_kernel<TT, false, false>, DuoSmem<TT>::BYTES)))
    if (T == 1) return PMS_DUO_INST(1);
    if (T == 2) return PMS_DUO_INST(2);
#undef PMS_DUO_INST
    return cudaErrorInvalidValue;
}

} // namespace

cudaError_t duo_kernel_prepare(bool comp, bool sens, int T, int *blocks_per_sm, size_t *smem_bytes)
{
    return dispatch(comp, sens, T, [&](auto kern, size_t smem) {
        *smem_bytes = smem;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kern, 96 * T + 32, smem);
    });
}

cudaError_t duo_kernel_launch(bool comp, bool sens, int T, const HsfmArgs &a, int n_worlds, size_t smem, cudaStream_t stream)
{
    return dispatch(comp, sens, T, [&](auto kern, size_t) {
        kern<<<n_worlds, 96 * T + 32, smem, stream>>>(a);
        return cudaGetLastError();
    });
}

} // namespace pms
