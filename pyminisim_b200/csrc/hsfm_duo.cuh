// hsfm_duo.cuh -- fused multi-step HSFM kernel with the step split over TWO KINDS OF WARPS per world
// ("pair warps" and "owner warps"), for batches that do not fill the GPU with one warp per world.
//
// Why: with one warp per world (hsfm_warp.cuh) a step is one dependent chain -- pair pass, then the per-pedestrian
// pass -- of about 2900 cycles (N = 32) that nothing hides when fewer than ~2 warps per scheduler are resident
// (C2: 1024 worlds = 1.7 warps per scheduler; C4 sharded over 8 GPUs: 512 worlds = 0.9).  The reference's explicit
// Euler step (_hsfm.py:290-294) moves the position with the OLD velocity, r' = r + v dt, so the positions of step
// t + 1 -- all the far-field pair force needs (_hsfm.py:124-126) -- are known as soon as the velocities of step t are.
// A world of T = ceil(N/32) tiles is served by 3T + 1 warps of one CTA:
//   * T OWNER warps: warp o keeps the complete state of the 32 pedestrians of tile o in registers.  Iteration t:
//     [A] body-frame dynamics of step t (_hsfm.py:69-110) with the forces F(r_t) the pair warps produced during iteration
//     t - 1 -> v_{t+1}, theta_{t+1}, omega_{t+1}; [B] r_{t+2} = r_{t+1} + v_{t+1} dt, waypoint tracker of step t + 1
//     (_hsfm.py:301-304) on it, and publication of r_{t+2} in the packed pair view.
//   * 2T PAIR warps: each evaluates ONE group of rotations of the unordered-pair scheme of hsfm_warp.cuh (one row tile
//     against one partner tile, 4 or 8 packed iterations) on the published r_{t+1} while the owners run [A] of step t, and
//     writes its partial forces (row side and reaction side) to shared memory; pair warp 0 also runs the rare body-contact
//     pass.
//   * ONE SERVICE warp: integrates the robot a block of 32 steps ahead (warp_robot_block) and evaluates the collision list
//     and the detector (core/_simulation.py:141-143, sensors/_pedestrian_detector.py:86-103) of every tile on r_{t+1}.
//     Measured as side duties of pair warps these two cost 410 and 500 cycles per step -- half a step -- so they get
//     their own warp, off everybody's critical path.
// All warps of a world meet at ONE named barrier per step; everything exchanged is double-buffered by step parity.  A step
// costs max(pair, owner) and every role is a short straight-line chain.
//
// Rare events: body contact (k_1, k_2 terms, _hsfm.py:127-132) needs the new velocities -- the pair warps only vote on
// the smallest distance; when a flag is up, pair warp 0 runs the contact pass of hsfm_warp.cuh after the barrier (one
// extra barrier).  The steady freeze of a non-looping tracker (_hsfm.py:302-304) stays inside the owner warp (position
// restored in [B], v / omega / theta in the following [A]).  Robot contact and |theta| >= 1e5 (Payne-Hanek sincos) are
// one early vote in the owner warp.
// Partial forces are added in a fixed order: deterministic, independent of the launch configuration; NOT bit-identical
// to hsfm_warp.cuh (different association of the same terms) -- compared with the oracle by tolerance.
// FP32 modes (compensated or plain state), N <= 64 (T = 1, 2), one world per CTA.
#pragma once
#include "hsfm_warp.cuh"
#include "hsfm_duo_api.cuh"

namespace pms {

constexpr int DUO_RING = 2 * WARP_ROBOT_BLOCK;      // robot slots: the block in use + the block being produced

template <int T> struct DuoSmem {
    static constexpr int Np = 32 * T;
    static constexpr int NU = 2 * T;                                    // units of the pair scheme (partial-force slots); duo_pair_warps(T) warps evaluate them
    // one FRAME per step parity holds what is double-buffered; every access is frame base + constant
    static constexpr size_t PV = 0;                                     // packed pair view [x | y][2 Np] float2
    static constexpr size_t XL = PV + 2 * 2 * Np * sizeof(float2);      // [xl | yl][Np] low parts of the position
    static constexpr size_t FP = XL + 2 * Np * sizeof(float);           // [NU units][row fx | row fy | reaction fx | reaction fy][32]
    static constexpr size_t FLAGS = FP + NU * 4 * 32 * sizeof(float);   // int4: contact vote of every pair warp
    static constexpr size_t FRAME = FLAGS + 16;
    static constexpr size_t VH = 2 * FRAME;                             // [vx | vy][Np] (contact pass)
    static constexpr size_t CF = VH + 2 * Np * sizeof(float);           // [cfx | cfy][Np] contact forces (zero between uses)
    static constexpr size_t ROBOT = (CF + 2 * Np * sizeof(float) + 15) & ~size_t(15);
    static constexpr size_t BYTES = (ROBOT + DUO_RING * sizeof(WarpRobotSlot) + sizeof(RobotRegs) + 15) & ~size_t(15);
};

// warp_sense without its early-out vote: in this kernel the sensor pass is a side duty of a pair warp and what counts is the
// length of its dependent chain (vote -> branch -> vote -> branch costs more than the 12 arithmetic instructions it skips)
__device__ __forceinline__ void duo_sense(const HsfmArgs &a, const WarpRobotSlot &rb, float qx, float qy, float px, float py,
                                          float pxl, float pyl, bool want_values, bool &det, bool &coll, double2 &vals)
{
    const float d2 = fmaf(qx, qx, qy * qy);
    const float d = d2 * fast_rsqrt(d2);
    const float u = fmaf(qx, rb.c, qy * rb.s) - a.sn_chf * d;          // d (cos(diff) - cos(fov/2))
    const float m = 1e-5f * d;
    det = d2 < a.sn_det_lo2 && u > m;                                   // sn_det_*: negative when the detector is off
    coll = d2 < a.sn_coll_lo2;
    const bool unsure = (!coll && d2 <= a.sn_coll_hi2) || (!det && d2 <= a.sn_det_hi2 && !(u < -m));
    const bool exact = unsure || (want_values && det);
    if (__any_sync(0xffffffffu, exact)) {                               // rare
        if (exact) {
            SenseOut so{false, false, 0.0, 0.0};
            const RobotSlot r64{rb.x, rb.y, rb.th, 0.0, 0.0};
            sense_ped_exact(a, r64, (double)px + (double)pxl, (double)py + (double)pyl, want_values, so);
            det = so.det; coll = so.coll; vals = make_double2(so.v0, so.v1);
        }
    }
}

// One unit of the unordered-pair scheme: the lane's row of tile A against partner tile B_, rotations K0 .. K0 + 2 NIT - 1
// (hsfm_warp.cuh: off-diagonal tiles take rotations 0..31, diagonal tiles 16..31 with rotation 16 on lanes < 16 only).
// fpart = [row fx | row fy | reaction fx | reaction fy][32]; a diagonal unit (A == B_) writes the sum to the row slots.
// Returns true when a pair of the unit is in body contact (d < r_ij).
template <int T, int A, int B_, int K0, int NIT>
__device__ __forceinline__ bool duo_unit(const HsfmArgs &a, const float2 *pl, int lane, float *fpart)
{
    constexpr int Np = 32 * T;
    const float ncexp = -a.kpp_f.cexp, rc = a.warp_rc;
    const float2 ncexp2 = make_float2(ncexp, ncexp), rc2 = make_float2(rc, rc);
    const float mx = pl[64 * A].x, my = pl[2 * Np + 64 * A].x;
    const float2 nmx = make_float2(-mx, -mx), nmy = make_float2(-my, -my);
    float2 fx = make_float2(0.f, 0.f), fy = fx, rx = fx, ry = fx;
    float d2min = CUDART_INF_F;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int off = 64 * B_ + K0 + 2 * it;
        float2 dx = __fadd2_rn(pl[off], nmx);
        const float2 dy = __fadd2_rn(pl[2 * Np + off], nmy);
        if (A == B_ && K0 == 16 && it == 0) dx.x = lane >= 16 ? 1e18f : dx.x;   // pedestrians l and l + 16 meet twice
        const float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
        float2 inv; inv.x = fast_rsqrt(d2.x); inv.y = fast_rsqrt(d2.y);
        const float2 arg = __ffma2_rn(d2, __fmul2_rn(inv, ncexp2), rc2);          // (r_ij - d) * log2(e) / B
        float2 e; e.x = fast_ex2(arg.x); e.y = fast_ex2(arg.y);
        const float2 c = __fmul2_rn(e, inv);                                       // e / d
        fx = __ffma2_rn(c, dx, fx); fy = __ffma2_rn(c, dy, fy);
        rx = __ffma2_rn(c, dx, rx); ry = __ffma2_rn(c, dy, ry);
        d2min = fminf(d2min, fminf(d2.x, d2.y));
        if (it != NIT - 1) {                                                       // partners advance two lanes
            const int src = (lane + 2) & 31;
            rx.x = __shfl_sync(0xffffffffu, rx.x, src); rx.y = __shfl_sync(0xffffffffu, rx.y, src);
            ry.x = __shfl_sync(0xffffffffu, ry.x, src); ry.y = __shfl_sync(0xffffffffu, ry.y, src);
        }
    }
    // hand the travelling accumulator to the lanes that own the partners (warp_deliver)
    constexpr int kx = K0 + 2 * (NIT - 1);
    const int below = (lane + 31) & 31;
    float sx = rx.x + __shfl_sync(0xffffffffu, rx.y, below);
    float sy = ry.x + __shfl_sync(0xffffffffu, ry.y, below);
    const int src = (lane - kx) & 31;
    sx = __shfl_sync(0xffffffffu, sx, src);
    sy = __shfl_sync(0xffffffffu, sy, src);
    const float A_ = a.warp_A;
    const float fxr = -A_ * (fx.x + fx.y), fyr = -A_ * (fy.x + fy.y);
    if constexpr (A == B_) {
        fpart[lane] = fmaf(A_, sx, fxr); fpart[32 + lane] = fmaf(A_, sy, fyr);
    } else {
        fpart[lane] = fxr; fpart[32 + lane] = fyr;
        fpart[64 + lane] = A_ * sx; fpart[96 + lane] = A_ * sy;
    }
    const float rij2 = a.kpp_f.rij * a.kpp_f.rij;
    return __any_sync(0xffffffffu, d2min < rij2);
}

// The two units of a one-tile world (rotations 16..23 and 24..31 of the diagonal tile) side by side in ONE warp: two independent
// chains, iteration by iteration; the lane's row force is the sum of the two units' (the value the owner used to add up).
template <int NIT>
__device__ __forceinline__ bool duo_unit_pair_t1(const HsfmArgs &a, const float2 *pl, int lane, float *fpart)
{
    constexpr int Np = 32, K0a = 16, K0b = 16 + 2 * NIT;
    const float ncexp = -a.kpp_f.cexp, rc = a.warp_rc;
    const float2 ncexp2 = make_float2(ncexp, ncexp), rc2 = make_float2(rc, rc);
    const float mx = pl[0].x, my = pl[2 * Np].x;
    const float2 nmx = make_float2(-mx, -mx), nmy = make_float2(-my, -my);
    const float2 z = make_float2(0.f, 0.f);
    float2 fxa = z, fya = z, rxa = z, rya = z, fxb = z, fyb = z, rxb = z, ryb = z;
    float d2min = CUDART_INF_F;
    auto iter = [&](int off, bool first, float2 &fx, float2 &fy, float2 &rx, float2 &ry) {
        float2 dx = __fadd2_rn(pl[off], nmx);
        const float2 dy = __fadd2_rn(pl[2 * Np + off], nmy);
        if (first) dx.x = lane >= 16 ? 1e18f : dx.x;                               // pedestrians l and l + 16 meet twice
        const float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
        float2 inv; inv.x = fast_rsqrt(d2.x); inv.y = fast_rsqrt(d2.y);
        const float2 arg = __ffma2_rn(d2, __fmul2_rn(inv, ncexp2), rc2);          // (r_ij - d) * log2(e) / B
        float2 e; e.x = fast_ex2(arg.x); e.y = fast_ex2(arg.y);
        const float2 c = __fmul2_rn(e, inv);                                       // e / d
        fx = __ffma2_rn(c, dx, fx); fy = __ffma2_rn(c, dy, fy);
        rx = __ffma2_rn(c, dx, rx); ry = __ffma2_rn(c, dy, ry);
        d2min = fminf(d2min, fminf(d2.x, d2.y));
    };
    auto travel = [&](float2 &rx, float2 &ry) {                                     // partners advance two lanes
        const int src = (lane + 2) & 31;
        rx.x = __shfl_sync(0xffffffffu, rx.x, src); rx.y = __shfl_sync(0xffffffffu, rx.y, src);
        ry.x = __shfl_sync(0xffffffffu, ry.x, src); ry.y = __shfl_sync(0xffffffffu, ry.y, src);
    };
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        iter(K0a + 2 * it, it == 0, fxa, fya, rxa, rya);
        iter(K0b + 2 * it, false, fxb, fyb, rxb, ryb);
        if (it != NIT - 1) { travel(rxa, rya); travel(rxb, ryb); }
    }
    const float A_ = a.warp_A;
    auto deliver = [&](int K0, float2 fx, float2 fy, float2 rx, float2 ry, float &ox, float &oy) {   // duo_unit's hand-over, diagonal unit
        const int kx = K0 + 2 * (NIT - 1);
        const int below = (lane + 31) & 31;
        float sx = rx.x + __shfl_sync(0xffffffffu, rx.y, below);
        float sy = ry.x + __shfl_sync(0xffffffffu, ry.y, below);
        const int src = (lane - kx) & 31;
        sx = __shfl_sync(0xffffffffu, sx, src);
        sy = __shfl_sync(0xffffffffu, sy, src);
        const float fxr = -A_ * (fx.x + fx.y), fyr = -A_ * (fy.x + fy.y);
        ox = fmaf(A_, sx, fxr); oy = fmaf(A_, sy, fyr);
    };
    float ax, ay, bx, by;
    deliver(K0a, fxa, fya, rxa, rya, ax, ay);
    deliver(K0b, fxb, fyb, rxb, ryb, bx, by);
    fpart[lane] = ax + bx; fpart[32 + lane] = ay + by;                             // = the sum the owner formed of the two units
    const float rij2 = a.kpp_f.rij * a.kpp_f.rij;
    return __any_sync(0xffffffffu, d2min < rij2);
}

// Two units of a two-tile world side by side in one warp: unit 1 = row tile A1 x partner tile B1 from rotation K1, unit 2 = A2 x B2
// from K2, NIT packed iterations each, separate accumulators (each unit's partial forces are bit-identical to duo_unit's), the
// iterations of the two chains alternating.  fp1 / fp2: the units' partial-force slots (layout of duo_unit).
template <int T, int A1, int B1, int K1, int A2, int B2, int K2, int NIT>
__device__ __forceinline__ bool duo_two_units(const HsfmArgs &a, const float2 *pl, int lane, float *fp1, float *fp2)
{
    constexpr int Np = 32 * T;
    const float ncexp = -a.kpp_f.cexp, rc = a.warp_rc;
    const float2 ncexp2 = make_float2(ncexp, ncexp), rc2 = make_float2(rc, rc);
    const float2 z = make_float2(0.f, 0.f);
    const float mx1 = pl[64 * A1].x, my1 = pl[2 * Np + 64 * A1].x, mx2 = pl[64 * A2].x, my2 = pl[2 * Np + 64 * A2].x;
    const float2 nmx1 = make_float2(-mx1, -mx1), nmy1 = make_float2(-my1, -my1), nmx2 = make_float2(-mx2, -mx2), nmy2 = make_float2(-my2, -my2);
    float2 fx1 = z, fy1 = z, rx1 = z, ry1 = z, fx2 = z, fy2 = z, rx2 = z, ry2 = z;
    float d2min = CUDART_INF_F;
    auto iter = [&](int off, bool mask16, float2 nmx, float2 nmy, float2 &fx, float2 &fy, float2 &rx, float2 &ry) {
        float2 dx = __fadd2_rn(pl[off], nmx);
        const float2 dy = __fadd2_rn(pl[2 * Np + off], nmy);
        if (mask16) dx.x = lane >= 16 ? 1e18f : dx.x;                              // pedestrians l and l + 16 meet twice
        const float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
        float2 inv; inv.x = fast_rsqrt(d2.x); inv.y = fast_rsqrt(d2.y);
        const float2 arg = __ffma2_rn(d2, __fmul2_rn(inv, ncexp2), rc2);          // (r_ij - d) * log2(e) / B
        float2 e; e.x = fast_ex2(arg.x); e.y = fast_ex2(arg.y);
        const float2 c = __fmul2_rn(e, inv);                                       // e / d
        fx = __ffma2_rn(c, dx, fx); fy = __ffma2_rn(c, dy, fy);
        rx = __ffma2_rn(c, dx, rx); ry = __ffma2_rn(c, dy, ry);
        d2min = fminf(d2min, fminf(d2.x, d2.y));
    };
    auto travel = [&](float2 &rx, float2 &ry) {                                     // partners advance two lanes
        const int src = (lane + 2) & 31;
        rx.x = __shfl_sync(0xffffffffu, rx.x, src); rx.y = __shfl_sync(0xffffffffu, rx.y, src);
        ry.x = __shfl_sync(0xffffffffu, ry.x, src); ry.y = __shfl_sync(0xffffffffu, ry.y, src);
    };
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        iter(64 * B1 + K1 + 2 * it, A1 == B1 && K1 == 16 && it == 0, nmx1, nmy1, fx1, fy1, rx1, ry1);
        iter(64 * B2 + K2 + 2 * it, A2 == B2 && K2 == 16 && it == 0, nmx2, nmy2, fx2, fy2, rx2, ry2);
        if (it != NIT - 1) { travel(rx1, ry1); travel(rx2, ry2); }
    }
    const float A_ = a.warp_A;
    auto deliver = [&](int K0, bool diag, float2 fx, float2 fy, float2 rx, float2 ry, float *fpart) {   // duo_unit's hand-over
        const int kx = K0 + 2 * (NIT - 1);
        const int below = (lane + 31) & 31;
        float sx = rx.x + __shfl_sync(0xffffffffu, rx.y, below);
        float sy = ry.x + __shfl_sync(0xffffffffu, ry.y, below);
        const int src = (lane - kx) & 31;
        sx = __shfl_sync(0xffffffffu, sx, src);
        sy = __shfl_sync(0xffffffffu, sy, src);
        const float fxr = -A_ * (fx.x + fx.y), fyr = -A_ * (fy.x + fy.y);
        if (diag) {
            fpart[lane] = fmaf(A_, sx, fxr); fpart[32 + lane] = fmaf(A_, sy, fyr);
        } else {
            fpart[lane] = fxr; fpart[32 + lane] = fyr;
            fpart[64 + lane] = A_ * sx; fpart[96 + lane] = A_ * sy;
        }
    };
    deliver(K1, A1 == B1, fx1, fy1, rx1, ry1, fp1);
    deliver(K2, A2 == B2, fx2, fy2, rx2, ry2, fp2);
    const float rij2 = a.kpp_f.rij * a.kpp_f.rij;
    return __any_sync(0xffffffffu, d2min < rij2);
}

// unit u of the world: T = 1: u0 = rotations 16..23, u1 = 24..31 of the one (diagonal) tile; T = 2: u0 = tile 0 diagonal,
// u1 / u2 = row tile 0 x partner tile 1 rotations 0..15 / 16..31, u3 = tile 1 diagonal
template <int T, int UPW>
__device__ __forceinline__ bool duo_far(const HsfmArgs &a, const float2 *pl, int u, int lane, float *fpart)
{
    if constexpr (T == 1) {
        if constexpr (UPW == 2) return duo_unit_pair_t1<4>(a, pl, lane, fpart);
        if (u == 0) return duo_unit<1, 0, 0, 16, 4>(a, pl, lane, fpart);
        return duo_unit<1, 0, 0, 24, 4>(a, pl, lane, fpart);
    } else if constexpr (UPW == 2) {
        // pair warp 0: units 0 + 1 (both row tile 0: diagonal, and tile 1 from rotation 0); pair warp 1: units 2 + 3 (row tile 0 and
        // row tile 1 against the same partners: tile 1, rotations 16..31).  `fpart` = slot of the warp's first unit, the second follows.
        if (u == 0) return duo_two_units<2, 0, 0, 16, 0, 1, 0, 8>(a, pl, lane, fpart, fpart + 128);
        return duo_two_units<2, 0, 1, 16, 1, 1, 16, 8>(a, pl, lane, fpart, fpart + 128);
    } else {
        if (u == 0) return duo_unit<2, 0, 0, 16, 8>(a, pl, lane, fpart);
        if (u == 1) return duo_unit<2, 0, 1, 0, 8>(a, pl, lane, fpart);
        if (u == 2) return duo_unit<2, 0, 1, 16, 8>(a, pl, lane, fpart);
        return duo_unit<2, 1, 1, 16, 8>(a, pl, lane, fpart);
    }
}

// body-contact terms of the whole world (every group; iterations without contact are skipped inside), by ONE warp
template <int T>
__device__ __forceinline__ void duo_contact(const HsfmArgs &a, const float2 *pl, const float *vxs, float *cf, int lane)
{
    if constexpr (T == 1) {
        warp_contact_pass(a, pl, vxs, cf, lane, 1, false, 0, 16, 4, 1);
        warp_contact_pass(a, pl, vxs, cf, lane, 1, false, 0, 24, 4, 1);
    } else {
        warp_contact_pass(a, pl, vxs, cf, lane, 2, false, 0, 16, 8, 1);
        warp_contact_pass(a, pl, vxs, cf, lane, 2, false, 1, 0, 8, 1);
        warp_contact_pass(a, pl, vxs, cf, lane, 2, false, 1, 16, 8, 2);
    }
}

// SENS: the instantiation that serves a sensor plan (see warp_robot_block in hsfm_warp.cuh); UPW: units per pair warp;
// MINB (two tiles, two units per warp): worlds per SM the kernel is compiled for instead of the default 5 (72 registers) --
// 7: 56 registers, a few more spills; 4: 96 registers; 3: 127 registers, no spills (hsfm_duo_api.cuh: duo_variant_blocks)
template <int T, bool COMP, bool SENS, int UPW, int MINB = 0>
__global__ void __launch_bounds__(duo_threads(T, UPW), MINB ? MINB : duo_min_blocks(T, UPW)) hsfm_duo_kernel(const __grid_constant__ HsfmArgs a)
{
    using L = DuoSmem<T>;
    constexpr int Np = 32 * T;
    constexpr int NU = duo_pair_warps(T, UPW);       // pair warps (DuoSmem::NU = 2T units, UPW per pair warp)
    constexpr int WPW = NU + T + 1;                  // warps of the world (= of the CTA): NU pair, T owner, 1 service
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int N = a.N;
    const int tid = threadIdx.x, lane = tid & 31;
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform (see hsfm_warp.cuh)
    const int w = blockIdx.x;
    const bool is_pair = wid < NU, is_service = wid == NU + T;
    const bool has_robot = a.robot_model != ROBOT_NONE;
    const int t_begin = 0, t_end = a.n_steps;

    unsigned char *base = smem_raw;
    auto frame = [&](int b) { return base + (size_t)b * L::FRAME; };
    auto pvb = [&](int b) { return reinterpret_cast<float2 *>(frame(b) + L::PV); };
    auto xlb = [&](int b) { return reinterpret_cast<float *>(frame(b) + L::XL); };
    auto fpb = [&](int b, int u) { return reinterpret_cast<float *>(frame(b) + L::FP) + u * 128; };
    auto flb = [&](int b) { return reinterpret_cast<int *>(frame(b) + L::FLAGS); };
    float *vh = reinterpret_cast<float *>(base + L::VH);
    float *cf = reinterpret_cast<float *>(base + L::CF);
    WarpRobotSlot *ring = reinterpret_cast<WarpRobotSlot *>(base + L::ROBOT);
    RobotRegs *rrp = reinterpret_cast<RobotRegs *>(base + L::ROBOT + DUO_RING * sizeof(WarpRobotSlot));

#ifdef PMS_DUO_TIMING      // experiment builds: busy / barrier cycles per warp, returned through the waypoint event log
    auto now = []() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; };
    long long tm_busy = 0, tm_wait = 0, tm_mark = now();
    auto world_sync = [&]() {
        const long long t0 = now();
        asm volatile("bar.sync 1, %0;" :: "n"(32 * WPW) : "memory");
        // BAR.SYNC.DEFER_BLOCKING lets the warp run on until its next shared-memory access: make one, and make the clock wait for it
        const int probe = *reinterpret_cast<volatile int *>(base + L::FLAGS);
        long long t1;
        asm volatile("{ .reg .b32 z; and.b32 z, %1, 0; cvt.u64.u32 %0, z; }" : "=l"(t1) : "r"(probe) : "memory");
        t1 += now();
        tm_busy += t0 - tm_mark; tm_wait += t1 - t0; tm_mark = t1;
    };
    auto timing_out = [&]() {
        if (lane == 0) {
            const int slot = w * WPW + wid;
            if (slot < a.ev_cap) { a.events[3 * slot] = (int)(tm_busy / a.n_steps); a.events[3 * slot + 1] = (int)(tm_wait / a.n_steps); a.events[3 * slot + 2] = wid; }
            atomicAdd(a.ev_count, 1);
        }
    };
#else
    auto world_sync = [&]() { asm volatile("bar.sync 1, %0;" :: "n"(32 * WPW) : "memory"); };
    auto timing_out = [&]() {};
#endif
    auto contact_flag = [&](int b) {
        const int4 f = *reinterpret_cast<const int4 *>(frame(b) + L::FLAGS);
        return NU == 1 ? f.x : (NU == 2 ? (f.x | f.y) : ((f.x | f.y) | (f.z | f.w)));
    };

    if (is_service) {
        // ================= service warp: robot a block ahead, collision list + detector of every tile =================
        const int words = (N + 31) >> 5;
        unsigned det_acc = 0, coll_acc = 0;
        if (has_robot) {
            if (lane == 0) robot_load(a, w, *rrp);
            __syncwarp();
            warp_robot_block<SENS>(a, w, t_begin, min(WARP_ROBOT_BLOCK, t_end - t_begin), rrp, ring, lane);   // SENS: + lidar / omni readings
        }
        world_sync();                                                   // #1
        world_sync();                                                   // #2
        for (int t = t_begin; t < t_end; ++t) {
            const int k = t - t_begin, b = k & 1, nb = b ^ 1;
            if (contact_flag(b)) world_sync();
            if (has_robot) {
                if ((k & (WARP_ROBOT_BLOCK - 1)) == 0 && t + WARP_ROBOT_BLOCK < t_end)   // (+ lidar / omni readings of the block's firing steps)
                    warp_robot_block<SENS>(a, w, t + WARP_ROBOT_BLOCK, min(WARP_ROBOT_BLOCK, t_end - t - WARP_ROBOT_BLOCK), rrp,
                                     ring + ((k + WARP_ROBOT_BLOCK) & (DUO_RING - 1)), lane);
                int dslot = -1, cslot = -1;          // sensor plan: records of the detector / collision list after this step
                if constexpr (SENS) {
                    if (a.sp.ped_on) {
                        if (a.sp.fire[SENSOR_DETECTOR]) dslot = a.sp.fire[SENSOR_DETECTOR][t];
                        if (a.sp.fire[SENSOR_COLLISIONS]) cslot = a.sp.fire[SENSOR_COLLISIONS][t];
                    }
                }
                const bool rec_vals = SENS && dslot >= 0 && a.sp.det_vals != nullptr;
                // ---- collision list + detector of step t: r_{t+1} against the robot after its step t; the tiles side by side
                //      (one vote for the rare exact path of all of them) ----
                const WarpRobotSlot &rb = ring[k & (DUO_RING - 1)];
                const bool last = (t == a.n_steps - 1);
                const float2 *p2 = pvb(nb);
                float px[T], py[T], pxl[T], pyl[T], d2[T], u[T], m[T];
                bool det[T], coll[T], exact[T];
                bool any_exact = false;
#pragma unroll
                for (int s_ = 0; s_ < T; ++s_) {
                    px[s_] = p2[64 * s_ + lane].x; py[s_] = p2[2 * Np + 64 * s_ + lane].x;
                    pxl[s_] = 0.f; pyl[s_] = 0.f;
                    if constexpr (COMP) { pxl[s_] = xlb(nb)[32 * s_ + lane]; pyl[s_] = xlb(nb)[Np + 32 * s_ + lane]; }
                    const float qx = (px[s_] - rb.xh) + (pxl[s_] - rb.xl), qy = (py[s_] - rb.yh) + (pyl[s_] - rb.yl);
                    d2[s_] = fmaf(qx, qx, qy * qy);
                    const float d = d2[s_] * fast_rsqrt(d2[s_]);
                    u[s_] = fmaf(qx, rb.c, qy * rb.s) - a.sn_chf * d;   // d (cos(diff) - cos(fov/2))
                    m[s_] = 1e-5f * d;
                    det[s_] = d2[s_] < a.sn_det_lo2 && u[s_] > m[s_];   // sn_det_*: negative when the detector is off
                    coll[s_] = d2[s_] < a.sn_coll_lo2;
                    const bool unsure = (!coll[s_] && d2[s_] <= a.sn_coll_hi2) || (!det[s_] && d2[s_] <= a.sn_det_hi2 && !(u[s_] < -m[s_]));
                    exact[s_] = unsure || ((last || rec_vals) && det[s_]);
                    any_exact = any_exact || exact[s_];
                }
                double2 dvals[T];
#pragma unroll
                for (int s_ = 0; s_ < T; ++s_) dvals[s_] = make_double2(0.0, 0.0);
                if (__any_sync(0xffffffffu, any_exact)) {               // rare: pedestrians inside a guard band, values on the last step
#pragma unroll
                    for (int s_ = 0; s_ < T; ++s_) {
                        if (exact[s_]) {
                            SenseOut so{false, false, 0.0, 0.0};
                            const RobotSlot r64{rb.x, rb.y, rb.th, 0.0, 0.0};
                            sense_ped_exact(a, r64, (double)px[s_] + (double)pxl[s_], (double)py[s_] + (double)pyl[s_], last || rec_vals, so);
                            det[s_] = so.det; coll[s_] = so.coll; dvals[s_] = make_double2(so.v0, so.v1);
                        }
                    }
                }
#pragma unroll
                for (int s_ = 0; s_ < T; ++s_) {
                    const unsigned dm = __ballot_sync(0xffffffffu, det[s_]);
                    const unsigned cm = __ballot_sync(0xffffffffu, coll[s_]);
                    det_acc += __popc(dm);
                    coll_acc += __popc(cm);
                    const int si = 32 * s_ + lane;
                    if (last) {
                        if (a.det_enabled && si < N) a.det_vals[(size_t)w * N + si] = dvals[s_];
                        if (lane == 0 && s_ < words) {
                            a.det_mask[(size_t)w * words + s_] = dm;
                            a.coll_mask[(size_t)w * words + s_] = cm;
                        }
                    }
                    if constexpr (SENS) {
                        if (rec_vals && si < N) a.sp.det_vals[((size_t)dslot * a.W + w) * N + si] = dvals[s_];
                        if (lane == 0 && s_ < words) {
                            if (dslot >= 0) a.sp.det_mask[((size_t)dslot * a.W + w) * words + s_] = dm;
                            if (cslot >= 0) a.sp.coll_mask[((size_t)cslot * a.W + w) * words + s_] = cm;
                        }
                    }
                }
            }
            world_sync();
        }
        timing_out();
        if (has_robot && lane == 0) {
            robot_store(a, w, *rrp);
            if (det_acc) atomicAdd(&a.det_count[w], (unsigned long long)det_acc);
            if (coll_acc) atomicAdd(&a.coll_count[w], (unsigned long long)coll_acc);
        }
    } else if (is_pair) {
        // ================= pair warp `wid`: one unit of the pair scheme =================
        const int u = wid;
        world_sync();                                                   // #1: r_t0 published by the owners
        {
            const bool touch = duo_far<T, UPW>(a, pvb(0) + lane, u, lane, fpb(0, UPW * u));
            if (lane == 0) flb(0)[u] = touch ? 1 : 0;
        }
        world_sync();                                                   // #2: F(r_t0), r_{t0+1}

        for (int t = t_begin; t < t_end; ++t) {
            const int k = t - t_begin, b = k & 1, nb = b ^ 1;
            if (contact_flag(b)) {                                      // body contact among r_t: needs v_t (published)
                if (u == 0) duo_contact<T>(a, pvb(b) + lane, vh, cf, lane);
                world_sync();
            }
            // ---- pair forces of step t + 1 on r_{t+1} ----
            if (t + 1 < t_end) {
                const bool touch = duo_far<T, UPW>(a, pvb(nb) + lane, u, lane, fpb(nb, UPW * u));
                if (lane == 0) flb(nb)[u] = touch ? 1 : 0;
            }
            world_sync();
        }
        timing_out();
    } else {
        // ================= owner warp: the complete state of one tile in registers =================
        const int tile = wid - NU;
        const int i = 32 * tile + lane;
        const bool valid = i < N;
        const size_t g = (size_t)w * N + (valid ? i : 0);
        const PedConsts<float> &kc = a.kc_f;
        const PairK<float> &kpr = a.kpr_f;
        // r_t | r_{t+1}; padding pedestrians: far away, pairwise distinct, at rest => every term that involves them is 0
        float x0 = 1e18f, y0 = 1e18f + 1e15f * (float)(i + 1), x0l = 0.f, y0l = 0.f;
        float vx = 0.f, vy = 0.f, vxl = 0.f, vyl = 0.f, th = 0.f, thl = 0.f, om = 0.f, oml = 0.f, c = 1.f, sn = 0.f, vm = 0.f;
        float wx0 = 0.f, wy0 = 0.f, nzx = 0.f, nzy = 0.f;
        int widx = 0, first_bad = 0;
        if (valid) {
            const float4 p4 = __ldcg(reinterpret_cast<const float4 *>(a.pos) + g);
            x0 = p4.x; y0 = p4.y; vx = p4.z; vy = p4.w;
            const float2 t2 = __ldcg(reinterpret_cast<const float2 *>(a.th) + g);
            th = t2.x; om = t2.y;
            const float2 q = __ldcg(reinterpret_cast<const float2 *>(a.wp) + g);
            wx0 = q.x; wy0 = q.y;
            widx = __ldcg(a.wp_index + g);
            vm = reinterpret_cast<const float *>(a.vmag)[g];
            fast_sincos(th, sn, c);
            if constexpr (COMP) {
                const float4 l4 = __ldcg(reinterpret_cast<const float4 *>(a.pos_lo) + g);
                const float2 tl = __ldcg(reinterpret_cast<const float2 *>(a.th_lo) + g);
                x0l = l4.x; y0l = l4.y; vxl = l4.z; vyl = l4.w; thl = tl.x; oml = tl.y;
                const float s0 = sn;
                sn = fmaf(c, thl, sn); c = fmaf(-s0, thl, c);
            }
            if (a.accel_noise) { nzx = (float)a.accel_noise[2 * g]; nzy = (float)a.accel_noise[2 * g + 1]; }
        }
        float x1 = x0, y1 = y0, x1l = x0l, y1l = y0l, wx1 = wx0, wy1 = wy0;
        bool frozen = false;                            // steady freeze decided by the tracker of the step in flight
        auto publish = [&](int b, float x, float y, float xl, float yl) {
            float2 *p2 = pvb(b);
            const int l0 = 64 * tile + lane, l1 = 64 * tile + ((lane + 31) & 31);
            p2[l0].x = x; p2[l0 + 32].x = x; p2[l1].y = x; p2[l1 + 32].y = x;
            p2[2 * Np + l0].x = y; p2[2 * Np + l0 + 32].x = y; p2[2 * Np + l1].y = y; p2[2 * Np + l1 + 32].y = y;
            if constexpr (COMP) { xlb(b)[i] = xl; xlb(b)[Np + i] = yl; }
        };
        // [B] of step s: r_{s+1} = r_s + v_s dt with the OLD velocity (_hsfm.py:290), tracker on the NEW position (:301),
        // steady freeze of the position (:302-304), publication for the pair warps
        auto advance_position = [&](int s, int b) {
            float nx = x1, ny = y1, nxl = x1l, nyl = y1l;
            if constexpr (COMP) {
                comp_add(nx, nxl, fmaf(vx, kc.dt, fmaf(vxl, kc.dt, x1l)));
                comp_add(ny, nyl, fmaf(vy, kc.dt, fmaf(vyl, kc.dt, y1l)));
            } else {
                nx = fmaf(vx, kc.dt, x1); ny = fmaf(vy, kc.dt, y1);
            }
            if (!valid) { nx = x1; ny = y1; nxl = x1l; nyl = y1l; }
            float nwx = wx1, nwy = wy1;
            const float ex = COMP ? (wx1 - nx) - nxl : wx1 - nx, ey = COMP ? (wy1 - ny) - nyl : wy1 - ny;
            const bool reached = valid && fmaf(ex, ex, ey * ey) <= kc.reach * kc.reach;
            bool steady = false;
            if (__any_sync(0xffffffffu, reached)) {
                if (reached) {
                    if (a.tracker_kind == TRACKER_FIXED) {       // _fixed_waypoint_tracker.py:68-87
                        int ni = widx + 1;
                        if (ni >= a.n_table) {
                            if (a.loop) ni = 0; else { ni = a.n_table - 1; steady = true; }
                        }
                        widx = ni;
                        const float2 nw = reinterpret_cast<const float2 *>(a.table)[g * a.n_table + ni];
                        nwx = nw.x; nwy = nw.y;
                        if (steady) { nx = x1; ny = y1; nxl = x1l; nyl = y1l; }
                    } else {                                     // _random_waypoint_tracker.py:66-91
                        const int cnt = a.q_count[g];
                        const int slot = atomicAdd(a.ev_count, 1);
                        if (cnt > 0) {
                            const int head = a.q_head[g];
                            const float2 nw = reinterpret_cast<const float2 *>(a.queue)[g * a.q_cap + head];
                            nwx = nw.x; nwy = nw.y;
                            a.q_head[g] = (head + 1) % a.q_cap;
                            a.q_count[g] = cnt - 1;
                            if (slot < a.ev_cap) { a.events[3 * slot] = s; a.events[3 * slot + 1] = w; a.events[3 * slot + 2] = i; }
                        } else if (slot < a.ev_cap) {            // starved queue: negative step marks it
                            a.events[3 * slot] = -1 - s; a.events[3 * slot + 1] = w; a.events[3 * slot + 2] = i;
                        }
                    }
                }
            }
            publish(b, nx, ny, nxl, nyl);
            x0 = x1; y0 = y1; x0l = x1l; y0l = y1l; wx0 = wx1; wy0 = wy1;
            x1 = nx; y1 = ny; x1l = nxl; y1l = nyl; wx1 = nwx; wy1 = nwy;
            frozen = steady;
        };

        publish(0, x0, y0, x0l, y0l);
        vh[i] = vx; vh[Np + i] = vy;
        cf[i] = 0.f; cf[Np + i] = 0.f;
        const bool robot_force = has_robot && a.robot_visible;
        const bool device_noise = !a.accel_noise && a.accel_noise_std > 0.0;
        world_sync();                                                   // #1
        advance_position(t_begin, 1);                                   // r_{t0+1}, tracker of step t0
        world_sync();                                                   // #2

        for (int t = t_begin; t < t_end; ++t) {
            const int k = t - t_begin, b = k & 1;
            const int cfl = contact_flag(b);
            if (cfl) world_sync();                                      // pair warp 0 ran the contact pass
            float ngx = nzx, ngy = nzy;
            if (device_noise) {                                         // uniform; fresh N(0, std) every step
                const double2 z = accel_noise_draw(a, g, a.steps_done + t);
                ngx = (float)z.x; ngy = (float)z.y;
            }
            // ---- early block: forces, new heading, robot far field, one vote for the rare cases ----
            float fx, fy;
            if constexpr (T == 1 && NU == 1) {
                fx = fpb(b, 0)[lane]; fy = fpb(b, 0)[32 + lane];            // the pair warp added its two units up
            } else if constexpr (T == 1) {
                fx = fpb(b, 0)[lane] + fpb(b, 1)[lane]; fy = fpb(b, 0)[32 + lane] + fpb(b, 1)[32 + lane];
            } else {
                if (tile == 0) {
                    fx = (fpb(b, 0)[lane] + fpb(b, 1)[lane]) + fpb(b, 2)[lane];
                    fy = (fpb(b, 0)[32 + lane] + fpb(b, 1)[32 + lane]) + fpb(b, 2)[32 + lane];
                } else {
                    fx = (fpb(b, 1)[64 + lane] + fpb(b, 2)[64 + lane]) + fpb(b, 3)[lane];
                    fy = (fpb(b, 1)[96 + lane] + fpb(b, 2)[96 + lane]) + fpb(b, 3)[32 + lane];
                }
            }
            if (cfl) { fx += cf[i]; fy += cf[Np + i]; cf[i] = 0.f; cf[Np + i] = 0.f; }
            float nth = th, nthl = thl;
            if constexpr (COMP) comp_add(nth, nthl, fmaf(om, kc.dt, fmaf(oml, kc.dt, thl)));
            else nth = fmaf(om, kc.dt, th);
            const bool big = !(fabsf(nth) < 1.0e5f);
            bool rare = big;
            float rinv = 0.f, rsg = 0.f, qx = 0.f, qy = 0.f, rvx = 0.f, rvy = 0.f;
            if (robot_force) {                                          // _hsfm.py:80-82
                const WarpRobotSlot &rb = ring[k & (DUO_RING - 1)];
                qx = (x0 - rb.xh) + (x0l - rb.xl); qy = (y0 - rb.yh) + (y0l - rb.yl);
                rvx = rb.vx; rvy = rb.vy;
                pair_far(kpr, qx, qy, false, rinv, rsg, fx, fy);
                rare = rare || rsg > 0.f;
            }
            float sn_slow = 0.f, c_slow = 1.f;
            if (__any_sync(0xffffffffu, rare)) {
                if (robot_force && rsg > 0.f) pair_contact(kpr, qx, qy, rvx - vx, rvy - vy, rinv, rsg, fx, fy);
                if (big) { const float2 r = sincos_slow(nth); sn_slow = r.x; c_slow = r.y; }
            }
            // ---- [A] fp32 dynamics of step t (_hsfm.py:69-71,86-110,146-158,290-294), straight-line ----
            const float ddx = COMP ? (wx0 - x0) - x0l : wx0 - x0, ddy = COMP ? (wy0 - y0) - y0l : wy0 - y0;
            const float n2 = fmaf(ddx, ddx, ddy * ddy);
            float ri = fast_rsqrt(n2);
            ri = ri * fmaf(-0.5f * n2 * ri, ri, 1.5f);                       // one Newton step
            const float sc = vm * ri;
            const float vdx = ddx * sc, vdy = ddy * sc;
            const float f0x = kc.m_over_tau * (COMP ? (vdx - vx) - vxl : vdx - vx);
            const float f0y = kc.m_over_tau * (COMP ? (vdy - vy) - vyl : vdy - vy);
            const float vo = fmaf(c, vy, -sn * vx);
            const float uf = fmaf(f0x + fx, c, (f0y + fy) * sn);
            const float uo = fmaf(kc.ko, fmaf(fy, c, -fx * sn), -kc.kd * vo);
            const float klf = kc.kl * fast_sqrt(fmaf(f0x, f0x, f0y * f0y));
            const bool vd0 = vdx == 0.f && vdy == 0.f;                       // speed 0: the reference's atan2(0, 0) = 0
            const float ux = vd0 ? 1.f : vdx, uy = vd0 ? 0.f : vdy;
            const float werr = fast_atan2(fmaf(sn, ux, -c * uy), fmaf(c, ux, sn * uy));
            const float komega = kc.one_alpha * fast_sqrt(klf * kc.inv_alpha);
            const float dom = fmaf(-klf, werr, -komega * om);                // u_theta / I  (I cancels)
            const float dvbx = fmaf(kc.inv_m, uf, ngx), dvby = fmaf(kc.inv_m, uo, ngy);
            float nom = om, noml = oml;
            if constexpr (COMP) comp_add(nom, noml, fmaf(dom, kc.dt, oml));
            else nom = fmaf(dom, kc.dt, om);
            float nsn, nc;
            fast_sincos_small(nth, nsn, nc);
            nsn = big ? sn_slow : nsn; nc = big ? c_slow : nc;
            if constexpr (COMP) {
                const float s0 = nsn;
                nsn = fmaf(nc, nthl, nsn); nc = fmaf(-s0, nthl, nc);         // rotation by the low part of theta
            }
            const float a_ = dvbx * kc.dt, b_ = dvby * kc.dt;
            float nvx = vx, nvy = vy, nvxl = vxl, nvyl = vyl;
            if constexpr (COMP) {
                comp_add(nvx, nvxl, fmaf(nc, a_, -nsn * b_) + vxl);
                comp_add(nvy, nvyl, fmaf(nsn, a_, nc * b_) + vyl);
            } else {
                nvx += fmaf(nc, a_, -nsn * b_);
                nvy += fmaf(nsn, a_, nc * b_);
            }
            // steady freeze decided by the tracker of this step (_hsfm.py:302-304): pose restored, velocity zeroed
            const bool hold = frozen || !valid;
            vx = hold ? 0.f : nvx; vy = hold ? 0.f : nvy; vxl = hold ? 0.f : nvxl; vyl = hold ? 0.f : nvyl;
            om = hold ? 0.f : nom; oml = hold ? 0.f : noml;
            th = frozen ? th : nth; thl = frozen ? thl : nthl; c = frozen ? c : nc; sn = frozen ? sn : nsn;
            vh[i] = vx; vh[Np + i] = vy;
            if (!first_bad && valid && !isfinite((vx + vy) + (om + x1 + y1))) first_bad = (int)(a.steps_done + t + 1);
            // ---- [B] position of step t + 1 (r_{t+2}), its tracker, publication ----
            if (t + 1 < t_end) advance_position(t + 1, b);
            world_sync();
        }
        // ---- write-back: state after step t_end - 1 = (r_{t_end} = x1, v, theta, omega, waypoint of step t_end) ----
        if (valid) {
            reinterpret_cast<float4 *>(a.pos)[g] = make_float4(x1, y1, vx, vy);
            reinterpret_cast<float2 *>(a.th)[g] = make_float2(th, om);
            if constexpr (COMP) {
                reinterpret_cast<float4 *>(a.pos_lo)[g] = make_float4(x1l, y1l, vxl, vyl);
                reinterpret_cast<float2 *>(a.th_lo)[g] = make_float2(thl, oml);
            }
            reinterpret_cast<float2 *>(a.wp)[g] = make_float2(wx1, wy1);
            a.wp_index[g] = widx;
        }
        if (first_bad) atomicMin(&a.nonfinite[w], first_bad);
        timing_out();
    }
}

} // namespace pms
