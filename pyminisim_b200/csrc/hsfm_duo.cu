// hsfm_duo.cu -- instantiations and launch dispatch of the pair-warp / owner-warp kernel (hsfm_duo.cuh).
#include "hsfm_duo.cuh"

namespace pms {
namespace {

// variant: see hsfm_duo_api.cuh
template <typename F> cudaError_t dispatch(bool comp, bool sens, int T, int variant, F &&f)
{
#define PMS_DUO_INST(TT, UU)                                                                                                  \
    (sens ? (comp ? f(hsfm_duo_kernel<TT, true, true, UU>, DuoSmem<TT>::BYTES, duo_threads(TT, UU))                           \
                  : f(hsfm_duo_kernel<TT, false, true, UU>, DuoSmem<TT>::BYTES, duo_threads(TT, UU)))                         \
          : (comp ? f(hsfm_duo_kernel<TT, true, false, UU>, DuoSmem<TT>::BYTES, duo_threads(TT, UU))                          \
                  : f(hsfm_duo_kernel<TT, false, false, UU>, DuoSmem<TT>::BYTES, duo_threads(TT, UU))))
    if (T == 1) return PMS_DUO_INST(1, 2);
#define PMS_DUO_MINB(MB)                                                                                                      \
    (comp ? f(hsfm_duo_kernel<2, true, false, 2, MB>, DuoSmem<2>::BYTES, duo_threads(2, 2))                                   \
          : f(hsfm_duo_kernel<2, false, false, 2, MB>, DuoSmem<2>::BYTES, duo_threads(2, 2)))
    if (T == 2 && variant >= 3 && !sens)
        return variant == 3 ? PMS_DUO_MINB(7) : (variant == 4 ? PMS_DUO_MINB(4) : PMS_DUO_MINB(3));
#undef PMS_DUO_MINB
    if (T == 2) return variant >= 2 ? PMS_DUO_INST(2, 2) : PMS_DUO_INST(2, 1);
#undef PMS_DUO_INST
    return cudaErrorInvalidValue;
}

} // namespace

cudaError_t duo_kernel_prepare(bool comp, bool sens, int T, int variant, int *blocks_per_sm, size_t *smem_bytes)
{
    return dispatch(comp, sens, T, variant, [&](auto kern, size_t smem, int threads) {
        *smem_bytes = smem;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kern, threads, smem);
    });
}

cudaError_t duo_kernel_launch(bool comp, bool sens, int T, int variant, const HsfmArgs &a, int n_worlds, size_t smem, cudaStream_t stream)
{
    return dispatch(comp, sens, T, variant, [&](auto kern, size_t, int threads) {
        kern<<<n_worlds, threads, smem, stream>>>(a);
        return cudaGetLastError();
    });
}

} // namespace pms
