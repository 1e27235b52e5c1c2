// hsfm_duo.cu -- instantiations and launch dispatch of the pair-warp / owner-warp kernel (hsfm_duo.cuh).
#include "hsfm_duo.cuh"

namespace pms {
namespace {

template <typename F> cudaError_t dispatch(bool comp, int T, F &&f)
{
    if (T == 1) return comp ? f(hsfm_duo_kernel<1, true>, DuoSmem<1>::BYTES) : f(hsfm_duo_kernel<1, false>, DuoSmem<1>::BYTES);
    if (T == 2) return comp ? f(hsfm_duo_kernel<2, true>, DuoSmem<2>::BYTES) : f(hsfm_duo_kernel<2, false>, DuoSmem<2>::BYTES);
    return cudaErrorInvalidValue;
}

} // namespace

cudaError_t duo_kernel_prepare(bool comp, int T, int *blocks_per_sm, size_t *smem_bytes)
{
    return dispatch(comp, T, [&](auto kern, size_t smem) {
        *smem_bytes = smem;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kern, 96 * T + 32, smem);
    });
}

cudaError_t duo_kernel_launch(bool comp, int T, const HsfmArgs &a, int n_worlds, size_t smem, cudaStream_t stream)
{
    return dispatch(comp, T, [&](auto kern, size_t) {
        kern<<<n_worlds, 96 * T + 32, smem, stream>>>(a);
        return cudaGetLastError();
    });
}

} // namespace pms
