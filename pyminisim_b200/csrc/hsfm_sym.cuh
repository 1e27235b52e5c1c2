// hsfm_sym.cuh -- fused multi-step HSFM kernel that evaluates every UNORDERED pedestrian pair once
// (Newton's third law), for fp32 pair arithmetic (FP32 and MIXED modes) and N <= 256.
//
// f(i<-j) = -f(j<-i) holds exactly in the reference (equal radii; n, t antisymmetric, the tangential
// velocity difference symmetric: _hsfm.py:124-132), so each pair needs one rsqrt + one ex2 instead of two.
//
// Work decomposition of a world with T = ceil(N/32) tiles of 32 pedestrians into T*T warp tasks of 16
// rotations each:
//   diagonal tile (a,a)        : rotations k = 1..16   (k = 16 only for lanes < 16: each pair once)
//   off-diagonal (a,b), a < b  : rotations k = 0..15 and k = 16..31 (two tasks)
// In rotation k lane l owns "row" pedestrian 32a + l and meets partner 32b + ((l + k) & 31).  The force
// goes to the lane's row accumulator; the REACTION goes to an accumulator that travels with the partner:
// it lives in the lane that currently meets that partner and moves two lanes per iteration with
// __shfl_sync (two rotations are processed per iteration as one packed-FP32 pair, FFMA2/FMUL2/FADD2).
// Partner coordinates come from shared memory as one conflict-free LDS.64 per coordinate: the pair view
// stores P2[m] = (x[m], x[m+1 within the tile]) so rotations k and k+1 arrive already packed.
// Every task writes its row / column partial forces to shared memory; owners add them in a fixed order
// (deterministic, unlike atomics).
//
// Everything else (owner phase, robot warp + ring, named per-world barriers, dynamic chunk scheduler) is
// shared with hsfm_small.cuh.
#pragma once
#include "hsfm_small.cuh"

namespace pms {

constexpr int SYM_MAX_TILES = 8;          // N <= 256
constexpr int SYM_MAX_WARPS = 16;         // pair warps per world

__host__ __device__ inline int sym_warps(int T) { return T * T < SYM_MAX_WARPS ? T * T : SYM_MAX_WARPS; }

template <typename ST, bool SPLIT>
struct SymSmem {
    // owner state (11 ST arrays + int index) | scalar pair view xs ys vxs vys (float) | packed pairs px2 py2
    // (float2) | MIXED: xlo ylo (float) + pxlo2 pylo2 (float2) | task table (int2) | row / col partials | robot
    __host__ __device__ static size_t off_widx(int Np) { return sizeof(ST) * 11 * Np; }
    __host__ __device__ static size_t off_scalar(int Np) { return off_widx(Np) + sizeof(int) * Np; }
    __host__ __device__ static size_t off_p2(int Np) { return off_scalar(Np) + 4 * sizeof(float) * Np; }
    __host__ __device__ static size_t off_lo(int Np) { return off_p2(Np) + 2 * sizeof(float2) * Np; }
    __host__ __device__ static size_t off_tasks(int Np) { return off_lo(Np) + (SPLIT ? (2 * sizeof(float) + 2 * sizeof(float2)) * Np : 0); }
    __host__ __device__ static size_t off_part(int Np, int T) { return off_tasks(Np) + sizeof(int2) * T * T; }
    __host__ __device__ static size_t off_robot(int Np, int T) { return (off_part(Np, T) + 2 * sizeof(float2) * 32 * T * T + 15) & ~size_t(15); }
    __host__ __device__ static size_t bytes(int Np, int T)
    {
        return (off_robot(Np, T) + ROBOT_RING * sizeof(RobotSlot) + sizeof(RobotRegs) + 15) & ~size_t(15);
    }
};

// task q -> (a, b, k0): diagonal tasks first, then two half-tasks per off-diagonal tile pair
__device__ __forceinline__ void sym_decode(int q, int T, int &ta, int &tb, int &k0)
{
    if (q < T) { ta = q; tb = q; k0 = 1; return; }
    int p = (q - T) >> 1;
    k0 = ((q - T) & 1) * 16;
    int a = 0;
    while (p >= T - 1 - a) { p -= T - 1 - a; ++a; }
    ta = a; tb = a + 1 + p;
}

template <typename ST, bool SPLIT>
__device__ __forceinline__ void sym_chunk(const HsfmArgs &a, const int group, const int t_begin, const int t_end)
{
    using ST2 = typename Vec2<ST>::type;
    using L = SymSmem<ST, SPLIT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int Np = a.Np, N = a.N, T = Np >> 5;
    const int Wp = sym_warps(T);
    const int per_world = Wp * 32;
    const int n_compute = a.wpc * per_world;
    const int tid = threadIdx.x, lane = tid & 31;
    const bool is_robot_warp = tid >= n_compute;
    const int wl = is_robot_warp ? (tid - n_compute) : tid / per_world;
    const int r = tid - wl * per_world;
    const int wq = r >> 5;                                    // warp index inside the world
    const int w = group * a.wpc + wl;
    const bool world_ok = (wl < a.wpc) && (w < a.W);
    const bool pair_warp = !is_robot_warp && world_ok;
    const int i = wq * 32 + lane;                             // owned pedestrian (owner warps: wq < T)
    const bool owner_warp = pair_warp && wq < T;
    const bool owner = owner_warp && i < N;
    const bool has_robot = a.robot_model != ROBOT_NONE;

    unsigned char *base = smem_raw + (size_t)(wl < a.wpc ? wl : 0) * L::bytes(Np, T);
    const OwnerView<ST> ow(base, Np, L::off_widx(Np));
    float *xs = reinterpret_cast<float *>(base + L::off_scalar(Np));
    float *ys = xs + Np, *vxs = xs + 2 * Np, *vys = xs + 3 * Np;
    float2 *px2 = reinterpret_cast<float2 *>(base + L::off_p2(Np));
    float2 *py2 = px2 + Np;
    float *xlo = reinterpret_cast<float *>(base + L::off_lo(Np));
    float *ylo = xlo + Np;
    float2 *pxlo2 = reinterpret_cast<float2 *>(ylo + Np);
    float2 *pylo2 = pxlo2 + Np;
    int2 *tasks = reinterpret_cast<int2 *>(base + L::off_tasks(Np));
    float2 *rowF = reinterpret_cast<float2 *>(base + L::off_part(Np, T));
    float2 *colF = rowF + 32 * T * T;
    RobotSlot *ring = reinterpret_cast<RobotSlot *>(base + L::off_robot(Np, T));
    RobotRegs &rr = *reinterpret_cast<RobotRegs *>(base + L::off_robot(Np, T) + ROBOT_RING * sizeof(RobotSlot));

    const PairK<float> &kpp = a.kpp_f;
    const PairK<float> &kpr = a.kpr_f;
    const size_t g = (size_t)w * N + i;
    const int nxt = (i & ~31) | ((i - 1) & 31);               // slot whose second half is x[i]

    // pair view: scalar copy + packed (x[m], x[m+1]) copy; MIXED adds the low parts of the fp64 position
    auto publish = [&](ST x, ST y, ST vx, ST vy) {
        const float xh = (float)x, yh = (float)y;
        xs[i] = xh; ys[i] = yh; vxs[i] = (float)vx; vys[i] = (float)vy;
        px2[i].x = xh; px2[nxt].y = xh; py2[i].x = yh; py2[nxt].y = yh;
        if constexpr (SPLIT) {
            const float xl = (float)(x - (double)xh), yl = (float)(y - (double)yh);
            xlo[i] = xl; ylo[i] = yl;
            pxlo2[i].x = xl; pxlo2[nxt].y = xl; pylo2[i].x = yl; pylo2[nxt].y = yl;
        }
    };

    if (owner) {
        ST x, y, vx, vy, c, sn;
        load_pv(a, g, x, y, vx, vy);
        const ST2 t = __ldcg(reinterpret_cast<const ST2 *>(a.th) + g);
        const ST2 q = __ldcg(reinterpret_cast<const ST2 *>(a.wp) + g);
        m_sincos(t.x, &sn, &c);
        ow.x[i] = x; ow.y[i] = y; ow.vx[i] = vx; ow.vy[i] = vy; ow.th[i] = t.x; ow.om[i] = t.y;
        ow.wx[i] = q.x; ow.wy[i] = q.y; ow.vm[i] = reinterpret_cast<const ST *>(a.vmag)[g]; ow.c[i] = c; ow.s[i] = sn;
        ow.widx[i] = __ldcg(a.wp_index + g);
        publish(x, y, vx, vy);
    } else if (owner_warp) {
        // padding pedestrians (N <= i < Np): far away, distinct, zero velocity => every term with them is 0
        publish((ST)1e18, (ST)(1e18 + 1e15 * (i + 1)), (ST)0, (ST)0);
    }
    if (pair_warp) {
        for (int q = r; q < T * T; q += per_world) {
            int ta, tb, k0;
            sym_decode(q, T, ta, tb, k0);
            tasks[q] = make_int2(ta | (tb << 8), k0);
        }
    }

    // ---- robot warp (identical protocol to hsfm_small.cuh) ----
    if (is_robot_warp) {
        const bool robot_lane = world_ok && has_robot;
        int filled = t_begin;
        auto fill = [&](int upto) {
            if (!robot_lane) return;
            for (; filled < upto; ++filled) {
                robot_step(a, w, filled, rr);
                ring[(filled - t_begin) & (ROBOT_RING - 1)] = RobotSlot{rr.x, rr.y, rr.th, rr.vx, rr.vy};
            }
        };
        if (robot_lane) robot_load(a, w, rr);
        fill(min(t_end, t_begin + ROBOT_BLOCK));
        for (int tb = t_begin; tb < t_end; tb += ROBOT_BLOCK) {
            cta_barrier();
            fill(min(t_end, tb + 2 * ROBOT_BLOCK));
        }
        cta_barrier();
        if (robot_lane) robot_store(a, w, rr);
        return;
    }

    unsigned det_acc = 0, coll_acc = 0;
    int first_bad = 0;
    const float2 rij2 = make_float2(kpp.rij, kpp.rij), cexp2 = make_float2(kpp.cexp, kpp.cexp), A2 = make_float2(kpp.A, kpp.A);

    for (int t = t_begin; t < t_end; ++t) {
        if (((t - t_begin) & (ROBOT_BLOCK - 1)) == 0) cta_barrier();
        if (pair_warp) world_barrier(wl, per_world);                    // A: state t (and the task table) visible
        const RobotSlot &rb = ring[(t - t_begin) & (ROBOT_RING - 1)];
        if (pair_warp) {
            for (int q = wq; q < T * T; q += Wp) {
                const int2 tk = tasks[q];
                const int ta = tk.x & 0xff, tb = tk.x >> 8, k0 = tk.y;
                const bool diag = ta == tb;
                const int ri = ta * 32 + lane;                          // row pedestrian of this lane
                const float mx = xs[ri], my = ys[ri];
                const float2 m2x = make_float2(mx, mx), m2y = make_float2(my, my);
                float2 l2x = make_float2(0.f, 0.f), l2y = l2x;
                if constexpr (SPLIT) { l2x = make_float2(xlo[ri], xlo[ri]); l2y = make_float2(ylo[ri], ylo[ri]); }
                float2 fx2 = make_float2(0.f, 0.f), fy2 = fx2;         // row force, two packed partial sums
                float2 rx2 = fx2, ry2 = fx2;                            // travelling reactions: .x rotation k, .y rotation k+1
                float cfx = 0.f, cfy = 0.f;                             // row contact force (rare path)
                const float2 *bx = px2 + tb * 32, *by = py2 + tb * 32;
#pragma unroll 2
                for (int k = k0; k < k0 + 16; k += 2) {
                    const int m = (lane + k) & 31;
                    const float2 X = bx[m], Y = by[m];                  // partners of rotations k and k+1
                    float2 dx = __fadd2_rn(m2x, make_float2(-X.x, -X.y));
                    float2 dy = __fadd2_rn(m2y, make_float2(-Y.x, -Y.y));
                    if constexpr (SPLIT) {
                        const float2 XL = pxlo2[tb * 32 + m], YL = pylo2[tb * 32 + m];
                        dx = __fadd2_rn(dx, __fadd2_rn(l2x, make_float2(-XL.x, -XL.y)));
                        dy = __fadd2_rn(dy, __fadd2_rn(l2y, make_float2(-YL.x, -YL.y)));
                    }
                    if (diag && k == 15) dx.y = lane >= 16 ? 1e18f : dx.y;   // rotation 16: each pair once
                    const float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
                    float2 inv; inv.x = fast_rsqrt(d2.x); inv.y = fast_rsqrt(d2.y);
                    const float2 d = __fmul2_rn(d2, inv);
                    const float2 sg = __fadd2_rn(rij2, make_float2(-d.x, -d.y));
                    const float2 arg = __fmul2_rn(sg, cexp2);
                    float2 e; e.x = fast_ex2(arg.x); e.y = fast_ex2(arg.y);
                    const float2 c = __fmul2_rn(e, __fmul2_rn(inv, A2));       // A e / d
                    fx2 = __ffma2_rn(c, dx, fx2); fy2 = __ffma2_rn(c, dy, fy2);
                    const float2 nc = make_float2(-c.x, -c.y);
                    rx2 = __ffma2_rn(nc, dx, rx2); ry2 = __ffma2_rn(nc, dy, ry2);
                    if (__any_sync(0xffffffffu, fmaxf(sg.x, sg.y) > 0.0f)) {   // body contact somewhere in the warp
                        const float mvx = vxs[ri], mvy = vys[ri];
                        const int j0 = tb * 32 + m, j1 = tb * 32 + ((m + 1) & 31);
                        float ax = 0.f, ay = 0.f, bxx = 0.f, byy = 0.f;
                        pair_contact(kpp, dx.x, dy.x, vxs[j0] - mvx, vys[j0] - mvy, inv.x, sg.x, ax, ay);
                        pair_contact(kpp, dx.y, dy.y, vxs[j1] - mvx, vys[j1] - mvy, inv.y, sg.y, bxx, byy);
                        cfx += ax + bxx; cfy += ay + byy;
                        rx2.x -= ax; ry2.x -= ay; rx2.y -= bxx; ry2.y -= byy;
                    }
                    if (k + 2 < k0 + 16) {                                     // partners advance two lanes
                        const int src = (lane + 2) & 31;
                        rx2.x = __shfl_sync(0xffffffffu, rx2.x, src); rx2.y = __shfl_sync(0xffffffffu, rx2.y, src);
                        ry2.x = __shfl_sync(0xffffffffu, ry2.x, src); ry2.y = __shfl_sync(0xffffffffu, ry2.y, src);
                    }
                }
                // .y holds the reaction for partner (lane + k0 + 15), .x for (lane + k0 + 14): the same partner's
                // two accumulators sit one lane apart
                const int below = (lane + 31) & 31;
                const float rx = rx2.x + __shfl_sync(0xffffffffu, rx2.y, below);
                const float ry = ry2.x + __shfl_sync(0xffffffffu, ry2.y, below);
                rowF[q * 32 + lane] = make_float2((fx2.x + fx2.y) + cfx, (fy2.x + fy2.y) + cfy);
                colF[q * 32 + ((lane + k0 + 14) & 31)] = make_float2(rx, ry);
            }
            world_barrier(wl, per_world);                               // B: partials visible, reads of state t done
        }
        if (owner_warp) {
            SenseOut so{false, false, 0.0, 0.0};
            const bool last = (t == a.n_steps - 1);
            if (owner) {
                float fx = 0.f, fy = 0.f;
                for (int q = 0; q < T * T; ++q) {                       // fixed order: deterministic
                    const int2 tk = tasks[q];
                    if ((tk.x & 0xff) == wq) { const float2 p = rowF[q * 32 + lane]; fx += p.x; fy += p.y; }
                    if ((tk.x >> 8) == wq) { const float2 p = colF[q * 32 + lane]; fx += p.x; fy += p.y; }
                }
                if (has_robot && a.robot_visible) {                     // robot term (_hsfm.py:80-82)
                    float dx, dy;
                    if constexpr (SPLIT) {
                        const float rxh = (float)rb.x, ryh = (float)rb.y;
                        dx = (xs[i] - rxh) + (xlo[i] - (float)(rb.x - (double)rxh));
                        dy = (ys[i] - ryh) + (ylo[i] - (float)(rb.y - (double)ryh));
                    } else { dx = xs[i] - (float)rb.x; dy = ys[i] - (float)rb.y; }
                    pair_accum(kpr, dx, dy, (float)rb.vx - vxs[i], (float)rb.vy - vys[i], false, fx, fy);
                }
                so = owner_update<ST>(a, ow, i, g, w, t, (ST)fx, (ST)fy, rb, has_robot, last, first_bad, publish);
            }
            if (has_robot) {
                const unsigned dm = __ballot_sync(0xffffffffu, so.det);
                const unsigned cm = __ballot_sync(0xffffffffu, so.coll);
                det_acc += __popc(dm);
                coll_acc += __popc(cm);
                if (last && lane == 0) {
                    const int words = (N + 31) >> 5;
                    if (wq < words) {
                        a.det_mask[(size_t)w * words + wq] = dm;
                        a.coll_mask[(size_t)w * words + wq] = cm;
                    }
                }
            }
        }
    }
    cta_barrier();

    if (owner) {
        store_pv(a, g, ow.x[i], ow.y[i], ow.vx[i], ow.vy[i]);
        ST2 t2; t2.x = ow.th[i]; t2.y = ow.om[i];
        reinterpret_cast<ST2 *>(a.th)[g] = t2;
        ST2 w2; w2.x = ow.wx[i]; w2.y = ow.wy[i];
        reinterpret_cast<ST2 *>(a.wp)[g] = w2;
        a.wp_index[g] = ow.widx[i];
        if (first_bad) atomicMin(&a.nonfinite[w], first_bad);
    }
    if (owner_warp && has_robot && lane == 0) {
        if (det_acc) atomicAdd(&a.det_count[w], (unsigned long long)det_acc);
        if (coll_acc) atomicAdd(&a.coll_count[w], (unsigned long long)coll_acc);
    }
}

template <typename ST, bool SPLIT>
__global__ void __launch_bounds__(544, 1) hsfm_sym_kernel(const __grid_constant__ HsfmArgs a)
{
    if (a.n_chunks <= 1) {
        sym_chunk<ST, SPLIT>(a, blockIdx.x, 0, a.n_steps);
        return;
    }
    __shared__ int s_item;
    const int n_groups = (a.W + a.wpc - 1) / a.wpc;
    const int total = n_groups * a.n_chunks;
    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(a.work_counter, 1);
        cta_barrier();
        const int item = s_item;
        cta_barrier();
        if (item >= total) break;
        const int chunk = item / n_groups, group = item - chunk * n_groups;
        if (chunk > 0) {
            if (threadIdx.x == 0) {
                int seen;
                do {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(a.group_done + group) : "memory");
                    if (seen < chunk) __nanosleep(200);
                } while (seen < chunk);
            }
            cta_barrier();
        }
        const int t_begin = chunk * a.chunk_steps;
        const int t_end = min(a.n_steps, t_begin + a.chunk_steps);
        sym_chunk<ST, SPLIT>(a, group, t_begin, t_end);
        __threadfence();
        cta_barrier();
        if (threadIdx.x == 0)
            asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(a.group_done + group), "r"(chunk + 1) : "memory");
    }
}

} // namespace pms
