// hsfm_large.cuh -- large single crowds (N > 992): placeholder until the cluster-tiled kernel lands.
#pragma once
#include "hsfm_small.cuh"
struct pms_launch_info_fwd;
namespace pms {
struct LargeScratch { void *buf = nullptr; };
inline void large_free(LargeScratch &s) { if (s.buf) cudaFree(s.buf); s.buf = nullptr; }
template <typename Info>
inline int launch_large(void *, LargeScratch &, HsfmArgs &, int, cudaStream_t, Info *) { return -4; }
} // namespace pms
