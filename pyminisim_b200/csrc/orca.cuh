// orca.cuh -- ORCA velocity-obstacle solve: placeholder until the kernel lands.
#pragma once
#include "pms_common.cuh"
namespace pms {
constexpr int ORCA_MAX_NEIGHBORS = 16;
struct OrcaState { void *buf = nullptr; };
inline void orca_free(OrcaState &s) { if (s.buf) cudaFree(s.buf); s.buf = nullptr; }
inline int orca_alloc(OrcaState &, int, int, double, double, int, double, double, double) { return 0; }
} // namespace pms
