// hsfm_large.cuh -- HSFM step for crowds that do not fit one CTA (N > 992, up to millions).
//
// One launch per simulation step (the step needs every pedestrian's state at time t, so the state is
// double-buffered in HBM and the launch boundary is the grid-wide barrier):
//   robot_window_kernel : one thread per world, fp64: the robot of every step of a window of LARGE_RESORT steps (it does not
//                         depend on the pedestrians), once per window
//   hsfm_large_kernel   : grid = worlds x i-tiles.  A CTA owns TI pedestrians (one per thread) and streams
//                         the whole world's "pair view" (structure-of-arrays x, y, vx, vy [, xlo, ylo])
//                         through shared memory in j-tiles of TJ with cp.async.bulk (TMA 1-D bulk copy)
//                         completing on an mbarrier, LARGE_STAGES buffers deep.  CTAs are launched in clusters; each
//                         CTA of a cluster fetches 1/C of every j-tile and MULTICASTS it into the shared
//                         memory of all C CTAs, so a tile is read from L2 once per cluster.  Buffer reuse is
//                         guarded by a split cluster barrier (arrive after compute, wait before refill).
// The pair math is the same packed-FP32 (FFMA2) code as the small-world kernel.
//
// Exact j-tile pruning (fp32 pair arithmetic): A exp((r_ij - d)/B) is computed as ex2.approx.ftz of an exponent that is
// below -126 for d > r_ij + 126 B / log2(e) (7.6 m with the default parameters), where the instruction returns exactly 0 and
// the pair contributes exactly nothing.  large_box_kernel keeps one bounding box per j-tile; a cluster skips every j-tile
// whose box is farther than that from the box of its own pedestrians -- no load, no arithmetic, bit-identical forces.
// The pruning needs spatially coherent indices, so the kernel works on a spatially sorted copy of the state: every
// LARGE_RESORT steps the pedestrians are sorted by (strip of one cut-off width, y) with a cub radix sort and
// position / velocity / heading are gathered into sorted buffers; per-pedestrian data that the host owns (waypoints,
// speeds, tracker tables and queues, detector outputs, event ids) stay in the caller's order and are reached through
// the permutation.  At the end of the call the state is scattered back.  Any input order gets the pruning.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include "hsfm_small.cuh"

namespace pms {

constexpr int LARGE_TI = 128;        // pedestrians (threads) per CTA
constexpr int LARGE_TJ = 512;        // pedestrians per streamed j-tile
constexpr int LARGE_STAGES = 4;      // j-tile buffers in shared memory (TMA runs LARGE_STAGES - 1 tiles ahead)

struct LargeArgs {
    // ping-pong state (element g = w*N + i); src is read, dst is written
    void *pos_src, *vel_src, *th_src, *pos_dst, *vel_dst, *th_dst;
    // FP32 with the compensated state (a.comp): vel_* hold the float4 low parts of {x,y,vx,vy}, thlo_* the float2 low
    // parts of {theta, omega}
    void *thlo_src, *thlo_dst;
    // pair view, structure-of-arrays of PT: arrays x, y, vx, vy (+ float xlo, ylo), each Npad long per world
    const void *view_src;
    void *view_dst;
    size_t view_stride;              // elements between consecutive arrays ( = W * Npad )
    int Npad;                        // N rounded up to a multiple of LARGE_TJ
    int tiles_i, tiles_j;
    int cluster;                     // CTAs per cluster (multicast width)
    int step;                        // global step index of this launch (events, controls)
    int last;                        // 1 if this is the last step of the pms_hsfm_step call
    const RobotSlot *robot_slot;     // (W) robot state for this step
    const float4 *boxes;             // (W, tiles_j) bounding box {xmin, xmax, ymin, ymax} of every j-tile of view_src, or NULL
    const float4 *chunk_boxes;       // (W, Npad / 32) the same per chunk of 32 pedestrians
    // chunk kernel: the boxes of view_dst are produced by the step itself (no box launch between two steps)
    float4 *boxes_dst;               // tile boxes of view_dst, merged with atomic min / max (reset to empty beforehand)
    float4 *boxes_reset;             // the third tile-box buffer: reset to empty for the step after the next
    float4 *chunk_boxes_dst;         // chunk boxes of view_dst
    int tiles_c;                     // chunk-kernel CTAs per world
    float cut2;                      // squared distance beyond which the pair force is exactly 0 (fp32 pair arithmetic)
    const int *perm;                 // (W*N) sorted slot -> caller's global pedestrian index, or NULL (identity)
};

struct LargeScratch {
    int sm_count = 148;                 // SMs of the device (set by the API from cudaDeviceProp): launch heuristics
    void *pos2 = nullptr, *vel2 = nullptr, *th2 = nullptr, *thlo2 = nullptr, *thlo3 = nullptr;
    void *view[2] = {nullptr, nullptr};
    RobotSlot *robot_slot = nullptr;
    RobotRegs *robot_regs = nullptr;
    float4 *boxes[3] = {nullptr, nullptr, nullptr}, *chunk_boxes[2] = {nullptr, nullptr};
    // spatially sorted working copy: sorted state (ping), keys / permutation double buffers, cub scratch
    void *pos3 = nullptr, *vel3 = nullptr, *th3 = nullptr;
    unsigned long long *keys[2] = {nullptr, nullptr};
    int *perm[2] = {nullptr, nullptr};
    void *sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    // the robots of a window are integrated on a side stream while the main stream sorts (fork / join events)
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int Npad = 0;
    // The sorted working copy outlives a call: while the caller's state arrays have not been written by the host
    // (pms_api.cu clears sorted_valid) and the last sort is less than LARGE_RESORT steps old, the next call goes on in it
    // -- a loop of one-step calls pays one sort per LARGE_RESORT steps, not one per call.
    bool sorted_valid = false;
    int since_sort = 0;              // steps taken in the current sorted copy
    int cur = 0;                     // which ping-pong buffer holds the newest sorted state (and its pair view)
    int bq = 0, cq = 0;              // chunk kernel: tile- / chunk-box buffers that hold the boxes of that state
    unsigned window = 0;             // windows so far (robot slot buffers alternate)
    bool view_valid = false;
    size_t view_bytes = 0;
};

inline void large_free(LargeScratch &s)
{
    void *p[] = {s.thlo2, s.thlo3, s.pos2, s.vel2, s.th2, s.view[0], s.view[1], s.robot_slot, s.robot_regs, s.boxes[0], s.boxes[1], s.boxes[2], s.chunk_boxes[0], s.chunk_boxes[1],
                 s.pos3, s.vel3, s.th3, s.keys[0], s.keys[1], s.perm[0], s.perm[1], s.sort_tmp};
    for (void *q : p) if (q) cudaFree(q);
    if (s.ev_fork) cudaEventDestroy(s.ev_fork);
    if (s.ev_join) cudaEventDestroy(s.ev_join);
    if (s.aux) cudaStreamDestroy(s.aux);
    s = LargeScratch{};
}

// ---- PTX helpers: mbarrier, bulk copy (TMA), cluster barrier ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_copy_multicast(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                                    uint16_t cta_mask)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// ---- robot: one thread per world ---------------------------------------------------------------------
__device__ __forceinline__ void large_robot_step(const HsfmArgs &a, int w, RobotRegs *regs, RobotSlot *slot, int step, int first, int last)
{
    if (w >= a.W || a.robot_model == ROBOT_NONE) return;
    RobotRegs r;
    if (first) robot_load(a, w, r); else r = regs[w];
    robot_step(a, w, step, r);
    regs[w] = r;
    slot[w] = RobotSlot{r.x, r.y, r.th, r.vx, r.vy};
    if (last) robot_store(a, w, r);
}
// The robot's motion does not depend on the pedestrians (controls + map veto only): the slots of steps t0 .. t1-1 are
// computed ahead, slot[(t - t0) * W + w].
__global__ void robot_window_kernel(const __grid_constant__ HsfmArgs a, RobotRegs *regs, RobotSlot *slot, int t0, int t1)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    for (int t = t0; t < t1; ++t)
        large_robot_step(a, w, regs, slot + (size_t)(t - t0) * a.W, t, t == 0, t == a.n_steps - 1);
}
// lidar / omni readings of the firing steps of the window (sensor plan): one warp per world, behind robot_window_kernel on
// the side stream
__global__ void sensor_window_kernel(const __grid_constant__ HsfmArgs a, const RobotSlot *slot, int t0, int t1)
{
    const int w = blockIdx.x, lane = threadIdx.x;
    for (int t = t0; t < t1; ++t) sensor_fires_block(a, w, t, 1, slot + (size_t)(t - t0) * a.W + w, lane);
}

// ---- pair view initialisation from the state (first step of a call) --------------------------------------
template <typename PT, typename ST, bool SPLIT>
__global__ void large_view_kernel(const __grid_constant__ HsfmArgs a, PT *view, size_t stride, int Npad)
{
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= (size_t)a.W * Npad) return;
    const int w = (int)(idx / Npad), i = (int)(idx - (size_t)w * Npad);
    PT x, y, vx = 0, vy = 0;
    float xl = 0.f, yl = 0.f;
    if (i < a.N) {
        ST sx, sy, svx, svy;
        load_pv(a, (size_t)w * a.N + i, sx, sy, svx, svy);
        if constexpr (SPLIT) {
            x = (float)sx; y = (float)sy; xl = (float)(sx - (double)x); yl = (float)(sy - (double)y);
        } else { x = (PT)sx; y = (PT)sy; }
        vx = (PT)svx; vy = (PT)svy;
    } else {   // padding: far away, contributes exactly 0
        x = (PT)1e18; y = (PT)(1e18 + 1e15 * ((i - a.N) % 1024 + 1));
    }
    view[idx] = x; view[stride + idx] = y; view[2 * stride + idx] = vx; view[3 * stride + idx] = vy;
    if constexpr (SPLIT) {
        float *lo = reinterpret_cast<float *>(view + 4 * stride);
        lo[idx] = xl; lo[stride + idx] = yl;
    }
}

// ---- bounding box of every j-tile of the pair view (exact pruning, see the file header) ------------------------
// One CTA of 128 threads per (world, j-tile): a box per chunk of 32 pedestrians (the warp-level skip unit of the step
// kernel) and their union per tile (the cluster-level skip unit).  Padding pedestrians are left out (an all-padding chunk
// has an empty box, which is far from everything); a chunk that holds a non-finite coordinate gets an infinite box (never
// skipped: its NaN must reach every other pedestrian like in the reference).
constexpr int LARGE_RESORT = 128;     // steps between two spatial sorts
// Strip width in units of the cut-off distance.  Wider than 1: the boxes of strip s and strip s +- 2 stay more than the
// cut-off apart while the crowd drifts between two sorts (measured on C3, 256 steps: 31.0 us per step at 1.0, 26.0 at 1.25).
constexpr float LARGE_STRIP_SCALE = 1.25f;

// Sort key = world : strip : y.  Strips are columns as wide as the cut-off distance; inside a strip the pedestrians are
// ordered by y, upwards in even strips and downwards in odd ones (a snake through the world).  32 consecutive pedestrians (a chunk) then fill a compact box of one strip width, 512 (a tile) a piece of
// one or two strips -- without the jumps a Z-order curve makes at its quadrant boundaries.  value = the global index.
template <typename ST>
__global__ void large_sortkey_kernel(const __grid_constant__ HsfmArgs a, float inv_strip, unsigned long long *keys, int *vals)
{
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= (size_t)a.W * a.N) return;
    ST x, y, vx, vy;
    load_pv(a, g, x, y, vx, vy);
    const float sx = fminf(fmaxf(floorf((float)x * inv_strip) + 32768.f, 0.f), 65535.f);   // NaN -> strip 0
    unsigned yb = __float_as_uint((float)y);
    yb = (yb & 0x80000000u) ? ~yb : (yb | 0x80000000u);                                     // order-preserving float bits
    if ((unsigned)sx & 1u) yb = ~yb;               // boustrophedon: odd strips run downwards, so a chunk or tile that
                                                   // straddles two strips stays compact (it turns the corner)
    keys[g] = ((unsigned long long)(g / a.N) << 48) | ((unsigned long long)(unsigned)sx << 32) | yb;
    vals[g] = (int)g;
}
// sorted <- caller order (gather) or caller order <- sorted (scatter) of position / velocity / heading
template <typename ST, bool GATHER>
__global__ void large_permute_kernel(size_t n, const int *perm, const void *pos_in, const void *vel_in, const void *th_in,
                                     void *pos_out, void *vel_out, void *th_out, const void *thlo_in = nullptr,
                                     void *thlo_out = nullptr)
{
    using ST2 = typename Vec2<ST>::type;
    const size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    const size_t src = GATHER ? (size_t)perm[k] : k, dst = GATHER ? k : (size_t)perm[k];
    if constexpr (sizeof(ST) == 4) {
        reinterpret_cast<float4 *>(pos_out)[dst] = reinterpret_cast<const float4 *>(pos_in)[src];
        if (thlo_in) {       // compensated state: the low parts travel with the pedestrian
            reinterpret_cast<float4 *>(vel_out)[dst] = reinterpret_cast<const float4 *>(vel_in)[src];
            reinterpret_cast<float2 *>(thlo_out)[dst] = reinterpret_cast<const float2 *>(thlo_in)[src];
        }
    } else {
        reinterpret_cast<double2 *>(pos_out)[dst] = reinterpret_cast<const double2 *>(pos_in)[src];
        reinterpret_cast<double2 *>(vel_out)[dst] = reinterpret_cast<const double2 *>(vel_in)[src];
    }
    reinterpret_cast<ST2 *>(th_out)[dst] = reinterpret_cast<const ST2 *>(th_in)[src];
}

constexpr int LARGE_CHUNK = 32;       // pedestrians per chunk box (one warp's rows / one warp-level skip unit)
constexpr int LARGE_HALF = 16;        // pedestrians per half-chunk box (the chunk kernel's skip unit on the partner side)

// Boxes of one chunk from the lanes' positions (empty lanes pass +-inf): the chunk box goes to out[ci], the boxes of its two
// halves (lanes 0-15 / 16-31) to out[n_chunks + 2 ci], out[n_chunks + 2 ci + 1] -- the half boxes live behind the chunk boxes in
// the same allocation.  A non-finite coordinate makes its boxes infinite (never skipped).  Returns the chunk box.
__device__ __forceinline__ float4 chunk_and_half_boxes(float xmin, float xmax, float ymin, float ymax, bool bad, int lane, size_t ci,
                                                       size_t n_chunks, float4 *out, float4 *out2)
{
    const float inf = CUDART_INF_F;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    const unsigned badm = __ballot_sync(0xffffffffu, bad);
    float4 hb = make_float4(xmin, xmax, ymin, ymax);
    if (badm & (lane < 16 ? 0x0000ffffu : 0xffff0000u)) hb = make_float4(-inf, inf, -inf, inf);
    xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, 16)); xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, 16));
    ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, 16)); ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, 16));
    float4 cb = make_float4(xmin, xmax, ymin, ymax);
    if (badm) cb = make_float4(-inf, inf, -inf, inf);
    if ((lane & 15) == 0) {
        out[n_chunks + 2 * ci + (lane >> 4)] = hb;
        if (out2) out2[n_chunks + 2 * ci + (lane >> 4)] = hb;
    }
    if (lane == 0) {
        out[ci] = cb;
        if (out2) out2[ci] = cb;
    }
    return cb;
}

template <typename PT>
__global__ void __launch_bounds__(128) large_box_kernel(const __grid_constant__ HsfmArgs a, int Npad, int tiles_j, const PT *view,
                                                        size_t stride, float4 *boxes, float4 *chunk_boxes,
                                                        float4 *chunk_boxes2, float4 *boxes_empty)
{
    // chunk_boxes2 (optional): second copy of the chunk boxes (the chunk kernel's destination buffer: its all-padding chunks
    // are never rewritten); boxes_empty (optional): tile-box buffer reset to empty boxes for the chunk kernel's atomics
    const int N = a.N;
    const int w = blockIdx.x / tiles_j, jt = blockIdx.x - w * tiles_j;
    const PT *xs = view + (size_t)w * Npad + (size_t)jt * LARGE_TJ, *ys = xs + stride;
    const float inf = CUDART_INF_F;
    float4 tb = make_float4(inf, -inf, inf, -inf);            // this warp's share of the tile box
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int m = 0; m < LARGE_TJ / 128; ++m) {                // warp q owns chunks q, q+4, q+8, ... of the tile
        const int k = m * 128 + threadIdx.x;
        float xmin = inf, xmax = -inf, ymin = inf, ymax = -inf;
        bool bad = false;
        if (jt * LARGE_TJ + k < N) {
            const float x = (float)xs[k], y = (float)ys[k];
            bad = !(fabsf(x) <= 3.0e38f) || !(fabsf(y) <= 3.0e38f);
            xmin = xmax = x; ymin = ymax = y;
        }
        const size_t ci = ((size_t)w * Npad + (size_t)jt * LARGE_TJ + k) / LARGE_CHUNK;     // this warp's chunk
        const float4 cb = chunk_and_half_boxes(xmin, xmax, ymin, ymax, bad, lane, ci, (size_t)a.W * Npad / LARGE_CHUNK, chunk_boxes, chunk_boxes2);
        tb.x = fminf(tb.x, cb.x); tb.y = fmaxf(tb.y, cb.y); tb.z = fminf(tb.z, cb.z); tb.w = fmaxf(tb.w, cb.w);
    }
    __shared__ float4 part[4];
    if (lane == 0) part[warp] = tb;
    __syncthreads();
    if (threadIdx.x == 0) {
        float4 bx = part[0];
        for (int q = 1; q < 4; ++q) {
            bx.x = fminf(bx.x, part[q].x); bx.y = fmaxf(bx.y, part[q].y); bx.z = fminf(bx.z, part[q].z); bx.w = fmaxf(bx.w, part[q].w);
        }
        boxes[blockIdx.x] = bx;
        if (boxes_empty) boxes_empty[blockIdx.x] = make_float4(inf, -inf, inf, -inf);
    }
}

// ---- the step kernel --------------------------------------------------------------------------------
// ---- robot term, O(1) dynamics, tracker, sensors of one pedestrian: shared by the tile-streaming kernel and the chunk kernel
// The per-pedestrian inputs of the owner step.  They are read in ONE batch (the L2 latencies overlap instead of queueing up
// behind the sincos / the robot term): by owner_load right before the step (tile-streaming kernel), or by cp.async into
// shared memory at the start of the chunk kernel, where they arrive while the pairs are evaluated.
template <typename ST>
struct OwnerIn {
    ST x, y, vx, vy, th, om, wpx, wpy, vm;
    float4 pl;      // FP32 compensated state: low parts of x y vx vy
    float2 tl;      //                          low parts of theta, omega
};
template <typename ST>
__device__ __forceinline__ OwnerIn<ST> owner_load(const HsfmArgs &a, const LargeArgs &la, size_t gs, size_t g)
{
    using ST2 = typename Vec2<ST>::type;
    OwnerIn<ST> in;
    in.pl = make_float4(0.f, 0.f, 0.f, 0.f); in.tl = make_float2(0.f, 0.f);
    if constexpr (sizeof(ST) == 4) {
        const float4 p = __ldcg(reinterpret_cast<const float4 *>(la.pos_src) + gs);
        in.x = p.x; in.y = p.y; in.vx = p.z; in.vy = p.w;
        if (a.comp) {
            in.pl = __ldcg(reinterpret_cast<const float4 *>(la.vel_src) + gs);
            in.tl = __ldcg(reinterpret_cast<const float2 *>(la.thlo_src) + gs);
        }
    } else {
        const double2 p = __ldcg(reinterpret_cast<const double2 *>(la.pos_src) + gs);
        const double2 v = __ldcg(reinterpret_cast<const double2 *>(la.vel_src) + gs);
        in.x = p.x; in.y = p.y; in.vx = v.x; in.vy = v.y;
    }
    const ST2 tw = __ldcg(reinterpret_cast<const ST2 *>(la.th_src) + gs);
    in.th = tw.x; in.om = tw.y;
    const ST2 wp = __ldcg(reinterpret_cast<const ST2 *>(a.wp) + g);
    in.wpx = wp.x; in.wpy = wp.y;
    in.vm = reinterpret_cast<const ST *>(a.vmag)[g];
    return in;
}

template <typename PT, typename ST, bool SPLIT>
__device__ __forceinline__ void large_owner(const HsfmArgs &a, const LargeArgs &la, int w, int i, bool valid, size_t gs, size_t g, int ie,
                                            int lane, PT mx, PT my, float mlx, float mly, PT mvx, PT mvy, PT fx, PT fy,
                                            const OwnerIn<ST> &in,
                                            PT &view_x, PT &view_y)            // out: the new position as the pair view holds it
{
    using ST2 = typename Vec2<ST>::type;
    const int N = a.N;
    const size_t vs = la.view_stride;
    const PairK<PT> &kpr = kpr_of(a, PT{});
    // ---- robot term, O(1) dynamics, tracker, sensors: identical to the small-world owner phase ----
    const bool has_robot = a.robot_model != ROBOT_NONE;
    RobotSlot rb{0, 0, 0, 0, 0};
    if (has_robot) rb = la.robot_slot[w];
    SenseOut so{false, false, 0.0, 0.0};
    int dslot = -1, cslot = -1;               // sensor plan: records of the detector / collision list after this step
    if (a.sp.ped_on) {
        if (a.sp.fire[SENSOR_DETECTOR]) dslot = a.sp.fire[SENSOR_DETECTOR][la.step];
        if (a.sp.fire[SENSOR_COLLISIONS]) cslot = a.sp.fire[SENSOR_COLLISIONS][la.step];
    }
    PT *vdst = reinterpret_cast<PT *>(la.view_dst) + (size_t)w * la.Npad;
    if (valid) {
        if (has_robot && a.robot_visible) {
            PT dx, dy;
            if constexpr (SPLIT) {
                const float rxh = (float)rb.x, ryh = (float)rb.y;
                dx = (mx - rxh) + (mlx - (float)(rb.x - (double)rxh));
                dy = (my - ryh) + (mly - (float)(rb.y - (double)ryh));
            } else { dx = mx - (PT)rb.x; dy = my - (PT)rb.y; }
            pair_accum(kpr, dx, dy, (PT)rb.vx - mvx, (PT)rb.vy - mvy, false, fx, fy);
        }
        const PedConsts<ST> &kc = kc_of(a, ST{});
        ST x = in.x, y = in.y, vx = in.vx, vy = in.vy;
        ST th = in.th, om = in.om, c, sn;
        m_sincos(th, &sn, &c);
        ST2 wp; wp.x = in.wpx; wp.y = in.wpy;
        const ST vm = in.vm;
        const ST x0 = x, y0 = y, th0 = th;
        ST nzx = 0, nzy = 0;
        if (a.accel_noise) { nzx = (ST)a.accel_noise[2 * g]; nzy = (ST)a.accel_noise[2 * g + 1]; }
        else if (a.accel_noise_std > 0.0) { const double2 z = accel_noise_draw(a, g, a.steps_done + la.step); nzx = (ST)z.x; nzy = (ST)z.y; }
        StateLo lo{0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, lo0 = lo;
        bool comp = false;
        if constexpr (sizeof(ST) == 4) comp = a.comp != 0;
        if constexpr (sizeof(ST) == 4) {
            if (comp) {
                const float4 pl = in.pl;
                const float2 tl = in.tl;
                lo = StateLo{pl.x, pl.y, pl.z, pl.w, tl.x, tl.y}; lo0 = lo;
                const float s0 = sn;
                sn = fmaf(c, tl.x, sn); c = fmaf(-s0, tl.x, c);
                ped_integrate_comp(kc, fx, fy, wp.x, wp.y, vm, x, y, vx, vy, th, om, c, sn, lo, nzx, nzy);
            } else {
                ped_integrate<ST>(kc, (ST)fx, (ST)fy, wp.x, wp.y, vm, x, y, vx, vy, th, om, c, sn, nzx, nzy);
            }
        } else {
            ped_integrate<ST>(kc, (ST)fx, (ST)fy, wp.x, wp.y, vm, x, y, vx, vy, th, om, c, sn, nzx, nzy);
        }
        const ST ex = (wp.x - x) - (ST)lo.x, ey = (wp.y - y) - (ST)lo.y;
        bool reached;
        if constexpr (sizeof(ST) == 8) reached = m_sqrt(ex * ex + ey * ey) <= kc.reach;
        else reached = ex * ex + ey * ey <= kc.reach * kc.reach;
        if (reached) {
            if (a.tracker_kind == TRACKER_FIXED) {
                int ni = a.wp_index[g] + 1;
                bool steady = false;
                if (ni >= a.n_table) { if (a.loop) ni = 0; else { ni = a.n_table - 1; steady = true; } }
                a.wp_index[g] = ni;
                wp = reinterpret_cast<const ST2 *>(a.table)[g * a.n_table + ni];
                if (steady) { x = x0; y = y0; th = th0; vx = 0; vy = 0; om = 0; lo = StateLo{lo0.x, lo0.y, 0.f, 0.f, lo0.th, 0.f}; }
            } else {
                const int cnt = a.q_count[g];
                const int slot = atomicAdd(a.ev_count, 1);
                if (cnt > 0) {
                    const int head = a.q_head[g];
                    wp = reinterpret_cast<const ST2 *>(a.queue)[g * a.q_cap + head];
                    a.q_head[g] = (head + 1) % a.q_cap;
                    a.q_count[g] = cnt - 1;
                    if (slot < a.ev_cap) { a.events[3 * slot] = la.step; a.events[3 * slot + 1] = w; a.events[3 * slot + 2] = ie; }
                } else if (slot < a.ev_cap) {
                    a.events[3 * slot] = -1 - la.step; a.events[3 * slot + 1] = w; a.events[3 * slot + 2] = ie;
                }
            }
            reinterpret_cast<ST2 *>(a.wp)[g] = wp;
        }
        // new state -> destination buffers + destination pair view
        if constexpr (sizeof(ST) == 4) {
            reinterpret_cast<float4 *>(la.pos_dst)[gs] = make_float4(x, y, vx, vy);
            if (comp) {
                reinterpret_cast<float4 *>(la.vel_dst)[gs] = make_float4(lo.x, lo.y, lo.vx, lo.vy);
                reinterpret_cast<float2 *>(la.thlo_dst)[gs] = make_float2(lo.th, lo.om);
            }
        } else {
            reinterpret_cast<double2 *>(la.pos_dst)[gs] = make_double2(x, y);
            reinterpret_cast<double2 *>(la.vel_dst)[gs] = make_double2(vx, vy);
        }
        ST2 t2; t2.x = th; t2.y = om;
        reinterpret_cast<ST2 *>(la.th_dst)[gs] = t2;
        if constexpr (SPLIT) {
            const float xh = (float)x, yh = (float)y;
            vdst[i] = xh; vdst[vs + i] = yh; vdst[2 * vs + i] = (float)vx; vdst[3 * vs + i] = (float)vy;
            vdst[4 * vs + i] = (float)(x - (double)xh); vdst[5 * vs + i] = (float)(y - (double)yh);
            view_x = xh; view_y = yh;
        } else {
            vdst[i] = (PT)x; vdst[vs + i] = (PT)y; vdst[2 * vs + i] = (PT)vx; vdst[3 * vs + i] = (PT)vy;
            view_x = (PT)x; view_y = (PT)y;
        }
        if (!m_finite(x + y + vx + vy + th + om)) atomicMin(&a.nonfinite[w], (int)(a.steps_done + la.step + 1));
        if (has_robot) {
            const bool want = la.last != 0 || (dslot >= 0 && a.sp.det_vals != nullptr);
            if constexpr (sizeof(ST) == 4) {
                if (comp) so = sense_ped<double>(a, rb, (double)x + (double)lo.x, (double)y + (double)lo.y, want);
                else so = sense_ped<ST>(a, rb, x, y, want);
            } else {
                so = sense_ped<ST>(a, rb, x, y, want);
            }
        }
        if (la.last && has_robot && a.det_enabled) a.det_vals[g] = make_double2(so.v0, so.v1);
        if (dslot >= 0 && has_robot && a.sp.det_vals) a.sp.det_vals[(size_t)dslot * a.W * N + g] = make_double2(so.v0, so.v1);
    }
    if (has_robot) {
        const unsigned dm = __ballot_sync(0xffffffffu, so.det);
        const unsigned cm = __ballot_sync(0xffffffffu, so.coll);
        const int words = (N + 31) >> 5;
        if (lane == 0) {
            if (dm) atomicAdd(&a.det_count[w], (unsigned long long)__popc(dm));
            if (cm) atomicAdd(&a.coll_count[w], (unsigned long long)__popc(cm));
            if (!la.perm && (i >> 5) < words) {
                if (la.last) {
                    a.det_mask[(size_t)w * words + (i >> 5)] = dm;
                    a.coll_mask[(size_t)w * words + (i >> 5)] = cm;
                }
                if (dslot >= 0) a.sp.det_mask[((size_t)dslot * a.W + w) * words + (i >> 5)] = dm;
                if (cslot >= 0) a.sp.coll_mask[((size_t)cslot * a.W + w) * words + (i >> 5)] = cm;
            }
        }
        if (la.perm) {          // sorted slots: set the caller's bit (the masks / records were cleared before the call)
            if (la.last) {
                if (so.det) atomicOr(&a.det_mask[(size_t)w * words + (ie >> 5)], 1u << (ie & 31));
                if (so.coll) atomicOr(&a.coll_mask[(size_t)w * words + (ie >> 5)], 1u << (ie & 31));
            }
            if (dslot >= 0 && so.det) atomicOr(&a.sp.det_mask[((size_t)dslot * a.W + w) * words + (ie >> 5)], 1u << (ie & 31));
            if (cslot >= 0 && so.coll) atomicOr(&a.sp.coll_mask[((size_t)cslot * a.W + w) * words + (ie >> 5)], 1u << (ie & 31));
        }
    }
}

template <typename PT, typename ST, bool SPLIT>
__global__ void __launch_bounds__(LARGE_TI, 4) hsfm_large_kernel(const __grid_constant__ HsfmArgs a, const __grid_constant__ LargeArgs la)
{
    using ST2 = typename Vec2<ST>::type;
    constexpr int NARR = SPLIT ? 6 : 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // [LARGE_STAGES][NARR arrays][TJ] of PT (lo arrays are float == PT in MIXED), then 2 mbarriers
    PT *tiles = reinterpret_cast<PT *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + LARGE_STAGES * NARR * LARGE_TJ * sizeof(PT));
    int *tile_list = reinterpret_cast<int *>(full + LARGE_STAGES);      // [tiles_j] j-tiles this cluster has to visit, then their count

    const int tid = threadIdx.x, lane = tid & 31;
    const int w = blockIdx.x / la.tiles_i, it = blockIdx.x - w * la.tiles_i;
    const int i = it * LARGE_TI + tid;
    const int N = a.N;
    const bool valid = i < N;
    const size_t gs = (size_t)w * N + (valid ? i : 0);                 // slot in the (spatially sorted) working state
    const size_t g = la.perm ? (size_t)la.perm[gs] : gs;               // the caller's index of this pedestrian
    const int ie = (int)(g - (size_t)w * N);
    const uint32_t crank = la.cluster > 1 ? cluster_ctarank() : 0;
    const PT *vsrc = reinterpret_cast<const PT *>(la.view_src) + (size_t)w * la.Npad;
    const size_t vs = la.view_stride;

    if (tid == 0) {
        for (int q = 0; q < LARGE_STAGES; ++q) mbar_init(&full[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        // j-tiles to visit (warp 0, 32 tiles per pass): every CTA of the cluster derives the same list from the same boxes
        const bool prune = la.boxes != nullptr && sizeof(PT) == 4;
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 *bx = prune ? la.boxes + (size_t)w * la.tiles_j : nullptr;
        if (prune) {
            const int it0 = it - (int)crank;                                   // first i-tile of the cluster
            const int jlo = it0 * LARGE_TI / LARGE_TJ;
            const int jhi = min(la.tiles_j - 1, ((it0 + la.cluster) * LARGE_TI - 1) / LARGE_TJ);
            c = bx[jlo];
            for (int q = jlo + 1; q <= jhi; ++q) {
                const float4 o = bx[q];
                c.x = fminf(c.x, o.x); c.y = fmaxf(c.y, o.y); c.z = fminf(c.z, o.z); c.w = fmaxf(c.w, o.w);
            }
        }
        int K = 0;
        for (int base = 0; base < la.tiles_j; base += 32) {
            const int jt = base + lane;
            bool keep = jt < la.tiles_j;
            if (keep && prune) {
                const float4 o = bx[jt];
                const float gx = fmaxf(0.f, fmaxf(o.x - c.y, c.x - o.y)), gy = fmaxf(0.f, fmaxf(o.z - c.w, c.z - o.w));
                keep = !(gx * gx + gy * gy > la.cut2);                          // NaN / infinite boxes are never skipped
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) tile_list[K + __popc(m & ((1u << lane) - 1u))] = jt;
            K += __popc(m);
        }
        if (lane == 0) tile_list[la.tiles_j] = K;
    }
    __syncthreads();
    cluster_arrive();
    cluster_wait();          // every CTA of the cluster has initialised its barriers

    constexpr uint32_t kTileBytes = NARR * LARGE_TJ * sizeof(PT);
    auto issue = [&](int k, int jt) {   // thread 0 only: tile jt into buffer k % LARGE_STAGES
        const int b = k % LARGE_STAGES;
        PT *dst = tiles + (size_t)b * NARR * LARGE_TJ;
        mbar_expect_tx(&full[b], kTileBytes);
        const int C = la.cluster;
        const int part = LARGE_TJ / C;                      // elements fetched by this CTA per array
        const uint16_t mask = (uint16_t)((1u << C) - 1u);
#pragma unroll
        for (int arr = 0; arr < NARR; ++arr) {
            const PT *src = vsrc + (size_t)arr * vs + (size_t)jt * LARGE_TJ + (size_t)crank * part;
            PT *d = dst + arr * LARGE_TJ + crank * part;
            if (C > 1) bulk_copy_multicast(d, src, part * sizeof(PT), &full[b], mask);
            else bulk_copy(d, src, part * sizeof(PT), &full[b]);
        }
    };

    // my own pedestrian (pair view of the source state)
    const int iv = valid ? i : 0;
    const PT mx = vsrc[iv], my = vsrc[vs + iv], mvx = vsrc[2 * vs + iv], mvy = vsrc[3 * vs + iv];
    float mlx = 0.f, mly = 0.f;
    if constexpr (SPLIT) { mlx = (float)vsrc[4 * vs + iv]; mly = (float)vsrc[5 * vs + iv]; }
    const PairK<PT> &kpp = kpp_of(a, PT{});
    const PairK<PT> &kpr = kpr_of(a, PT{});

    PT fx = 0, fy = 0;
    float2 fxa = make_float2(0.f, 0.f), fya = fxa, fxb = fxa, fyb = fxa;
    const int i0 = i & ~31;
    const bool prune_chunks = la.boxes != nullptr && sizeof(PT) == 4;
    float4 wbox = make_float4(0.f, 0.f, 0.f, 0.f);               // box of this warp's 32 pedestrians
    if (prune_chunks) wbox = __ldg(la.chunk_boxes + ((size_t)w * la.Npad + i0) / LARGE_CHUNK);
    const int T = tile_list[la.tiles_j];
    if (tid == 0)
        for (int q = 0; q < LARGE_STAGES - 1 && q < T; ++q) issue(q, tile_list[q]);
    for (int k = 0; k < T; ++k) {
        const int jt = tile_list[k];
        if (k >= 1) cluster_wait();                  // everybody is done with tile k - 1: its buffer may be refilled
        if (k + LARGE_STAGES - 1 < T && tid == 0) issue(k + LARGE_STAGES - 1, tile_list[k + LARGE_STAGES - 1]);
        mbar_wait(&full[k % LARGE_STAGES], (k / LARGE_STAGES) & 1);
        const PT *xs = tiles + (size_t)(k % LARGE_STAGES) * NARR * LARGE_TJ;
        const PT *ys = xs + LARGE_TJ, *vxs = xs + 2 * LARGE_TJ, *vys = xs + 3 * LARGE_TJ;
        const int jbase = jt * LARGE_TJ;
        if constexpr (sizeof(PT) == 4) {
            const float *xlo = reinterpret_cast<const float *>(xs + 4 * LARGE_TJ), *ylo = xlo + LARGE_TJ;
            const float2 m2x = make_float2(mx, mx), m2y = make_float2(my, my);
            const float2 l2x = make_float2(mlx, mlx), l2y = make_float2(mly, mly);
            auto deltas = [&](int jl, auto masked, float2 &dxa, float2 &dxb, float2 &dya, float2 &dyb) {
                const float4 X = *reinterpret_cast<const float4 *>(xs + jl);
                const float4 Y = *reinterpret_cast<const float4 *>(ys + jl);
                dxa = __fadd2_rn(m2x, make_float2(-X.x, -X.y)); dxb = __fadd2_rn(m2x, make_float2(-X.z, -X.w));
                dya = __fadd2_rn(m2y, make_float2(-Y.x, -Y.y)); dyb = __fadd2_rn(m2y, make_float2(-Y.z, -Y.w));
                if constexpr (SPLIT) {
                    const float4 XL = *reinterpret_cast<const float4 *>(xlo + jl);
                    const float4 YL = *reinterpret_cast<const float4 *>(ylo + jl);
                    dxa = __fadd2_rn(dxa, __fadd2_rn(l2x, make_float2(-XL.x, -XL.y)));
                    dxb = __fadd2_rn(dxb, __fadd2_rn(l2x, make_float2(-XL.z, -XL.w)));
                    dya = __fadd2_rn(dya, __fadd2_rn(l2y, make_float2(-YL.x, -YL.y)));
                    dyb = __fadd2_rn(dyb, __fadd2_rn(l2y, make_float2(-YL.z, -YL.w)));
                }
                if constexpr (decltype(masked)::value) {
                    const int k = i - (jbase + jl);
                    dxa.x = k == 0 ? 1e18f : dxa.x; dxa.y = k == 1 ? 1e18f : dxa.y;
                    dxb.x = k == 2 ? 1e18f : dxb.x; dxb.y = k == 3 ? 1e18f : dxb.y;
                }
            };
            // one chunk of 32 partners: far-field pass in straight-line code (no vote between the 8 groups), one vote on the
            // largest overlap, then the rare contact pass with 1 / d and the overlap re-derived bit for bit (like hsfm_chunk_kernel)
            auto chunk = [&](int jc, auto masked) {
                float smax = -CUDART_INF_F;
#pragma unroll
                for (int jl = jc; jl < jc + LARGE_CHUNK; jl += 4) {
                    float2 dxa, dxb, dya, dyb, inva, invb, sa, sb;
                    deltas(jl, masked, dxa, dxb, dya, dyb);
                    pair_far2(kpp, dxa, dya, inva, sa, fxa, fya);
                    pair_far2(kpp, dxb, dyb, invb, sb, fxb, fyb);
                    smax = fmaxf(smax, fmaxf(fmaxf(sa.x, sa.y), fmaxf(sb.x, sb.y)));
                }
                if (!__any_sync(0xffffffffu, smax > 0.0f)) return;
                for (int jl = jc; jl < jc + LARGE_CHUNK; jl += 4) {
                    float2 dxa, dxb, dya, dyb, inva, invb, sa, sb;
                    deltas(jl, masked, dxa, dxb, dya, dyb);
                    pair_geom2(kpp, dxa, dya, inva, sa);
                    pair_geom2(kpp, dxb, dyb, invb, sb);
                    const float s4 = fmaxf(fmaxf(sa.x, sa.y), fmaxf(sb.x, sb.y));
                    if (!__any_sync(0xffffffffu, s4 > 0.0f)) continue;
                    const float4 VX = *reinterpret_cast<const float4 *>(vxs + jl);
                    const float4 VY = *reinterpret_cast<const float4 *>(vys + jl);
                    pair_contact(kpp, dxa.x, dya.x, VX.x - mvx, VY.x - mvy, inva.x, sa.x, fx, fy);
                    pair_contact(kpp, dxa.y, dya.y, VX.y - mvx, VY.y - mvy, inva.y, sa.y, fx, fy);
                    pair_contact(kpp, dxb.x, dyb.x, VX.z - mvx, VY.z - mvy, invb.x, sb.x, fx, fy);
                    pair_contact(kpp, dxb.y, dyb.y, VX.w - mvx, VY.w - mvy, invb.y, sb.y, fx, fy);
                }
            };
            // chunk by chunk (32 consecutive j): skipped when its box is out of reach of this warp's box (warp-uniform, exact);
            // the chunk [i0, i0+32) holds j == i
            unsigned near = (1u << (LARGE_TJ / LARGE_CHUNK)) - 1u;           // bit c: chunk c of this tile has to be visited
            if (prune_chunks) {                                                 // lane c tests chunk c
                bool keep = false;
                if (lane < LARGE_TJ / LARGE_CHUNK) {
                    const float4 o = __ldg(la.chunk_boxes + ((size_t)w * la.Npad + jbase) / LARGE_CHUNK + lane);
                    const float gx = fmaxf(0.f, fmaxf(o.x - wbox.y, wbox.x - o.y)), gy = fmaxf(0.f, fmaxf(o.z - wbox.w, wbox.z - o.w));
                    keep = !(gx * gx + gy * gy > la.cut2);
                }
                near = __ballot_sync(0xffffffffu, keep);
            }
            while (near) {
                const int jc = (__ffs(near) - 1) * LARGE_CHUNK;
                near &= near - 1;
                if (jbase + jc == i0) chunk(jc, std::true_type{});
                else chunk(jc, std::false_type{});
            }
        } else {
            const int jn = min(LARGE_TJ, N - jbase);
            for (int jl = 0; jl < jn; ++jl)
                pair_accum(kpp, mx - xs[jl], my - ys[jl], vxs[jl] - mvx, vys[jl] - mvy, jbase + jl == i, fx, fy);
        }
        cluster_arrive();                            // done reading the buffer of tile k
    }
    cluster_wait();
    if constexpr (sizeof(PT) == 4) {
        fx += (fxa.x + fxa.y) + (fxb.x + fxb.y);
        fy += (fya.x + fya.y) + (fyb.x + fyb.y);
    }

    PT view_x = 0, view_y = 0;
    OwnerIn<ST> in{};
    if (valid) in = owner_load<ST>(a, la, gs, g);
    large_owner<PT, ST, SPLIT>(a, la, w, i, valid, gs, g, ie, lane, mx, my, mlx, mly, mvx, mvy, fx, fy, in, view_x, view_y);
}

// ---- pruned regime: one warp per chunk, no cluster, no block barrier (FP32 pair view) ---------------------------------
// With exact pruning on the sorted copy a pedestrian meets ~200 of the 65 536 others, and the tile-streaming kernel above
// spends its time at the per-tile cluster barriers (ncu: 28 % issue, barrier + membar stalls).  Here every warp works alone:
// it owns the 32 pedestrians of one chunk, finds the chunks within reach of its own box (tile boxes 32 per ballot, then the 16
// chunk boxes of a near tile), gathers their indices in a per-warp list and walks it with a two-deep software pipeline -- the
// coalesced loads of chunk k + 1 (x, y, vx, vy of pedestrian `lane`, 512 B from L2) are in flight while chunk k, staged in the
// warp's own 512-byte shared-memory buffer, is evaluated with the broadcast LDS.128 / packed-FP32 pair code of the kernel above.
// Same pairs, same arithmetic per pair; the order of summation is (tile, chunk, j) like above, so the result is bit-identical
// to the tile-streaming kernel's.
// The step also produces what the next step's pruning needs -- the box of the warp's 32 new positions (chunk box) and, merged
// with atomic min / max, the box of their j-tile -- so that one launch per step remains.  min / max are exact in any order:
// the boxes equal large_box_kernel's bit for bit.
constexpr int CHUNK_LIST = 128;       // near half-chunks gathered per pass of the list
constexpr int CHUNK_TI = 64;          // KS = 1: threads per CTA, two independent warps (small CTAs spread evenly over 148 SMs)
// KS > 1 (j-split): KS warps share one chunk of 32 pedestrians -- a 65 536-pedestrian world is only 2048 warps (3.5 per
// scheduler), too few to hide the MUFU / L2 latency of the pair chains.  Warp q takes the near chunks at list positions
// q, q + KS, ...; the partial sums meet in shared memory and are added in warp order by warp 0, which then runs the
// per-pedestrian step.  Deterministic, but the order of summation differs from KS = 1 (same pairs, same per-pair arithmetic).
// KS is chosen by the number of chunks (launch_large_t): 1 from about 1200 chunks (38 000 pedestrians) upwards.

// atomic min / max of finite-or-infinite floats through their ordered integer images (+0.0f is added so that -0 counts as +0)
__device__ __forceinline__ void atomic_min_float(float *p, float v)
{
    v += 0.0f;
    if (v >= 0.f) atomicMin(reinterpret_cast<int *>(p), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned *>(p), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float *p, float v)
{
    v += 0.0f;
    if (v >= 0.f) atomicMax(reinterpret_cast<int *>(p), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned *>(p), __float_as_uint(v));
}

#ifndef PMS_CHUNK_MINB
#define PMS_CHUNK_MINB 8              // resident 64-thread CTAs per SM the register budget is set for (16 warps per SM)
#endif
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst_smem)), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_u32(dst_smem)), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async4(void *dst_smem, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(dst_smem)), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#ifndef PMS_CHUNK_MINB_HI
#define PMS_CHUNK_MINB_HI 14          // the same for crowds of several waves: 28 warps per SM at 72 registers
#endif
// MINB: register budget as resident 64-thread CTAs per SM.  One wave of warps (C3: 2048 warps, 13.8 per SM) runs fastest with
// 128 registers (23.1 vs 24.1 us per step at 80); several waves (1024 worlds x 512 pedestrians: 16 384 warps) gain 9-14 % from
// more resident warps (us per step at 128 / 80 / 72 / 64 registers: 1024 x 512: 93.4 / 84.9 / 83.0 / 81.5; 4096 x 512: 355 / 316 /
// 306 / 294; 256 x 992: 47.3 / 45.1 / 43.6 / 44.7).  launch_large_t chooses by the chunk count.
template <int KS, int CPB, int MINB = PMS_CHUNK_MINB>       // warps per chunk, chunks per CTA, register budget
__global__ void __launch_bounds__(32 * KS * CPB, MINB * 2 / (KS * CPB)) hsfm_chunk_kernel(const __grid_constant__ HsfmArgs a, const __grid_constant__ LargeArgs la)
{
    using PT = float;
    constexpr int NW = KS * CPB;
    __shared__ __align__(16) float stage[NW][2][4 * LARGE_CHUNK];
    __shared__ int near_list[NW][CHUNK_LIST];
    __shared__ float2 part[KS > 1 ? NW : 1][32];
    __shared__ __align__(16) float4 own16[CPB][2][32];                 // owner inputs: {x y vx vy}, their low parts
    __shared__ __align__(8) float2 own8[CPB][3][32];                   //               {theta omega}, low parts, waypoint
    __shared__ float own4[CPB][32];                                    //               speed
    __shared__ unsigned short near_tiles[NW][128];                     // near tiles of the current group of 128

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int q = wid % KS;                                            // this warp's share of the near list
    const int w = blockIdx.x / la.tiles_c, it = blockIdx.x - w * la.tiles_c;
    const int i = (it * CPB + wid / KS) * LARGE_CHUNK + lane;
    const int N = a.N;
    const bool valid = i < N;
    const size_t gs = (size_t)w * N + (valid ? i : 0);
    const size_t g = la.perm ? (size_t)la.perm[gs] : gs;
    const int ie = (int)(g - (size_t)w * N);
    const PT *vsrc = reinterpret_cast<const PT *>(la.view_src) + (size_t)w * la.Npad;
    const size_t vs = la.view_stride;
    const int iv = valid ? i : 0;
    const PT mx = vsrc[iv], my = vsrc[vs + iv], mvx = vsrc[2 * vs + iv], mvy = vsrc[3 * vs + iv];
    const PairK<PT> &kpp = a.kpp_f;
    const int i0 = i & ~31;
    constexpr int CPT = LARGE_TJ / LARGE_CHUNK;                        // chunks per tile
    const float4 *tile_box = la.boxes + (size_t)w * la.tiles_j;
    const float4 *chunk_box = la.chunk_boxes + (size_t)w * la.Npad / LARGE_CHUNK;
    const float4 wbox = __ldg(chunk_box + i0 / LARGE_CHUNK);          // box of this warp's 32 pedestrians
    int *list = near_list[wid];
    // the per-pedestrian step's inputs travel into shared memory (cp.async, no registers held) while the pairs are evaluated;
    // every thread reads back only what it copied itself
    const int oc = wid / KS;                                           // chunk of this CTA
    if (q == 0 && valid) {
        cp_async16(&own16[oc][0][lane], reinterpret_cast<const float4 *>(la.pos_src) + gs);
        cp_async8(&own8[oc][0][lane], reinterpret_cast<const float2 *>(la.th_src) + gs);
        if (a.comp) {
            cp_async16(&own16[oc][1][lane], reinterpret_cast<const float4 *>(la.vel_src) + gs);
            cp_async8(&own8[oc][1][lane], reinterpret_cast<const float2 *>(la.thlo_src) + gs);
        }
        cp_async8(&own8[oc][2][lane], reinterpret_cast<const float2 *>(a.wp) + g);
        cp_async4(&own4[oc][lane], reinterpret_cast<const float *>(a.vmag) + g);
    }

    PT fx = 0, fy = 0;
    float2 fxa = make_float2(0.f, 0.f), fya = fxa, fxb = fxa, fyb = fxa;
    const float2 m2x = make_float2(mx, mx), m2y = make_float2(my, my);

    auto near = [&](const float4 o) {                                 // NaN / infinite boxes are never skipped
        const float gx = fmaxf(0.f, fmaxf(o.x - wbox.y, wbox.x - o.y)), gy = fmaxf(0.f, fmaxf(o.z - wbox.w, wbox.z - o.w));
        return !(gx * gx + gy * gy > la.cut2);
    };
    // list entries are HALF-chunks (16 pedestrians); a stage holds two of them: entry k in slots 0-15, entry k + 1 in 16-31
    auto load_pair = [&](int k, int K) {
        const int e = k + (lane >> 4);
        if (e >= K) return make_float4(1e18f, 1e18f, 0.f, 0.f);        // odd tail: a partner far away contributes exactly 0
        const int j = list[e] * LARGE_HALF + (lane & 15);
        return make_float4(__ldcg(vsrc + j), __ldcg(vsrc + vs + j), __ldcg(vsrc + 2 * vs + j), __ldcg(vsrc + 3 * vs + j));
    };
    const int my_half = i0 / LARGE_HALF + (lane >> 4);                 // the half-chunk this lane's own pedestrian sits in
    // (x_i - x_j) of the 4 partners at slots jl .. jl + 3; self = slot of j == i in the stage, -1000 if it is not staged, or -1 on
    // every lane (uniform) when neither of the warp's own halves is staged
    auto deltas = [&](const float *xs, int jl, int self, float2 &dxa, float2 &dxb, float2 &dya, float2 &dyb) {
        const float4 X = *reinterpret_cast<const float4 *>(xs + jl);
        const float4 Y = *reinterpret_cast<const float4 *>(xs + LARGE_CHUNK + jl);
        dxa = __fadd2_rn(m2x, make_float2(-X.x, -X.y)); dxb = __fadd2_rn(m2x, make_float2(-X.z, -X.w));
        dya = __fadd2_rn(m2y, make_float2(-Y.x, -Y.y)); dyb = __fadd2_rn(m2y, make_float2(-Y.z, -Y.w));
        if (self != -1) {
            const int k = self - jl;
            dxa.x = k == 0 ? 1e18f : dxa.x; dxa.y = k == 1 ? 1e18f : dxa.y;
            dxb.x = k == 2 ? 1e18f : dxb.x; dxb.y = k == 3 ? 1e18f : dxb.y;
        }
    };
    // far-field term of a stage in straight-line code (no vote, no branch between the 8 groups: their MUFU chains overlap);
    // the largest overlap r_ij - d of the stage tells whether the rare body-contact pass has anything to do
    auto stage_far = [&](const float *xs, int self) {
        float smax = -CUDART_INF_F;
#pragma unroll
        for (int jl = 0; jl < LARGE_CHUNK; jl += 4) {
            float2 dxa, dxb, dya, dyb, inva, invb, sa, sb;
            deltas(xs, jl, self, dxa, dxb, dya, dyb);
            pair_far2(kpp, dxa, dya, inva, sa, fxa, fya);
            pair_far2(kpp, dxb, dyb, invb, sb, fxb, fyb);
            smax = fmaxf(smax, fmaxf(fmaxf(sa.x, sa.y), fmaxf(sb.x, sb.y)));
        }
        return smax;
    };
    // body-contact terms of a stage: 1 / d and the overlap are re-derived with the same operations (bit for bit), groups and
    // partners in the same order as the far pass
    auto stage_contact = [&](const float *xs, int self) {
        for (int jl = 0; jl < LARGE_CHUNK; jl += 4) {
            float2 dxa, dxb, dya, dyb, inva, invb, sa, sb;
            deltas(xs, jl, self, dxa, dxb, dya, dyb);
            pair_geom2(kpp, dxa, dya, inva, sa);
            pair_geom2(kpp, dxb, dyb, invb, sb);
            const float smax = fmaxf(fmaxf(sa.x, sa.y), fmaxf(sb.x, sb.y));
            if (!__any_sync(0xffffffffu, smax > 0.0f)) continue;
            const float4 VX = *reinterpret_cast<const float4 *>(xs + 2 * LARGE_CHUNK + jl);
            const float4 VY = *reinterpret_cast<const float4 *>(xs + 3 * LARGE_CHUNK + jl);
            pair_contact(kpp, dxa.x, dya.x, VX.x - mvx, VY.x - mvy, inva.x, sa.x, fx, fy);
            pair_contact(kpp, dxa.y, dya.y, VX.y - mvx, VY.y - mvy, inva.y, sa.y, fx, fy);
            pair_contact(kpp, dxb.x, dyb.x, VX.z - mvx, VY.z - mvy, invb.x, sb.x, fx, fy);
            pair_contact(kpp, dxb.y, dyb.y, VX.w - mvx, VY.w - mvy, invb.y, sb.y, fx, fy);
        }
    };
    auto flush = [&](int K) {                                         // evaluate the K chunks of the list
        __syncwarp();
        if (K == 0) return;
        float4 nxt = load_pair(0, K);
        for (int k = 0; k < K; k += 2) {
            float *buf = stage[wid][(k >> 1) & 1];
            buf[lane] = nxt.x; buf[LARGE_CHUNK + lane] = nxt.y; buf[2 * LARGE_CHUNK + lane] = nxt.z; buf[3 * LARGE_CHUNK + lane] = nxt.w;
            __syncwarp();                   // pair k staged; everybody is done with pair k - 2 (the other buffer may be refilled)
            const int ea = list[k], eb = k + 1 < K ? list[k + 1] : -1;
            if (k + 2 < K) nxt = load_pair(k + 2, K);
            // own pedestrian among the staged partners?  (uniform test on the warp's two halves, then the lane's own slot)
            const int h0 = i0 / LARGE_HALF;
            int self = -1;
            if (ea == h0 || ea == h0 + 1 || eb == h0 || eb == h0 + 1)
                self = ea == my_half ? (lane & 15) : (eb == my_half ? LARGE_HALF + (lane & 15) : -1000);
            const float smax = stage_far(buf, self);
            if (__any_sync(0xffffffffu, smax > 0.0f)) stage_contact(buf, self);
        }
        __syncwarp();
    };

    // The scan is a chain of dependent L2 reads (tile boxes, then the chunk boxes of the near tiles), so the reads are issued in
    // batches: 128 tile boxes at a time, then the chunk boxes of up to 8 near tiles (two tiles per warp-wide load, 16 lanes each).
    static_assert(CPT * 2 == 32, "one tile = 32 half-chunks per warp-wide box load");
    const float4 *half_box = la.chunk_boxes + (size_t)a.W * la.Npad / LARGE_CHUNK + (size_t)w * la.Npad / LARGE_HALF;
    int K = 0, P = 0;                                                  // entries in this warp's list; near chunks seen so far
    for (int base0 = 0; base0 < la.tiles_j; base0 += 128) {
        float4 tb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int jt = base0 + 32 * u + lane;
            tb[u] = jt < la.tiles_j ? __ldg(tile_box + jt) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        unsigned tm[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) tm[u] = __ballot_sync(0xffffffffu, base0 + 32 * u + lane < la.tiles_j && near(tb[u]));
        int cnt = 0;                                                   // near tiles of this group, ascending, in shared memory
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if ((tm[u] >> lane) & 1u) near_tiles[wid][cnt + __popc(tm[u] & ((1u << lane) - 1u))] = (unsigned short)(32 * u + lane);
            cnt += __popc(tm[u]);
        }
        __syncwarp();
        for (int r = 0; r < cnt; r += 4) {                             // near tiles of rank r .. r + 3: lane = half-chunk of the tile
            float4 cbx[4];
            int cid[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                cid[v] = r + v < cnt ? (base0 + (int)near_tiles[wid][r + v]) * (2 * CPT) + lane : -1;
                cbx[v] = cid[v] >= 0 ? __ldg(half_box + cid[v]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                if (r + v >= cnt) break;                               // uniform
                bool keep = cid[v] >= 0 && near(cbx[v]);
                unsigned mc = __ballot_sync(0xffffffffu, keep);
                if constexpr (KS > 1) {                                // position P + rank in the common list decides the owner warp
                    const int pos = P + __popc(mc & ((1u << lane) - 1u));
                    P += __popc(mc);
                    keep = keep && (pos % KS == q);
                    mc = __ballot_sync(0xffffffffu, keep);
                }
                const int n = __popc(mc);
                if (K + n > CHUNK_LIST) { flush(K); K = 0; }
                if (keep) list[K + __popc(mc & ((1u << lane) - 1u))] = cid[v];
                K += n;
            }
        }
    }
    flush(K);
    fx += (fxa.x + fxa.y) + (fxb.x + fxb.y);
    fy += (fya.x + fya.y) + (fyb.x + fyb.y);
    if constexpr (KS > 1) {
        part[wid][lane] = make_float2(fx, fy);
        __syncthreads();
        if (q != 0) return;
        fx = part[wid][lane].x; fy = part[wid][lane].y;
#pragma unroll
        for (int r = 1; r < KS; ++r) { fx += part[wid + r][lane].x; fy += part[wid + r][lane].y; }
    }
    float nx = 0.f, ny = 0.f;
    OwnerIn<float> in{};
    cp_async_wait_all();
    if (valid) {
        const float4 p = own16[oc][0][lane];
        const float2 tw = own8[oc][0][lane], wp = own8[oc][2][lane];
        in.x = p.x; in.y = p.y; in.vx = p.z; in.vy = p.w; in.th = tw.x; in.om = tw.y; in.wpx = wp.x; in.wpy = wp.y; in.vm = own4[oc][lane];
        in.pl = make_float4(0.f, 0.f, 0.f, 0.f); in.tl = make_float2(0.f, 0.f);
        if (a.comp) { in.pl = own16[oc][1][lane]; in.tl = own8[oc][1][lane]; }
    }
    large_owner<float, float, false>(a, la, w, i, valid, gs, g, ie, lane, mx, my, 0.f, 0.f, mvx, mvy, fx, fy, in, nx, ny);

    // ---- boxes of the new positions (same rules as large_box_kernel) ----
    {
        const size_t gt = (blockIdx.x * (size_t)CPB + wid / KS) * LARGE_CHUNK + lane;
        if (gt < (size_t)a.W * la.tiles_j) la.boxes_reset[gt] = make_float4(CUDART_INF_F, -CUDART_INF_F, CUDART_INF_F, -CUDART_INF_F);
        const float inf = CUDART_INF_F;
        float xmin = inf, xmax = -inf, ymin = inf, ymax = -inf;
        bool bad = false;
        if (valid) {
            bad = !(fabsf(nx) <= 3.0e38f) || !(fabsf(ny) <= 3.0e38f);
            xmin = xmax = nx; ymin = ymax = ny;
        }
        if (i0 >= N) return;                                      // an all-padding chunk keeps the empty boxes of the box kernel
        const float4 cb = chunk_and_half_boxes(xmin, xmax, ymin, ymax, bad, lane, ((size_t)w * la.Npad + i0) / LARGE_CHUNK,
                                               (size_t)a.W * la.Npad / LARGE_CHUNK, la.chunk_boxes_dst, nullptr);
        if (lane < 4) {
            float *tb = reinterpret_cast<float *>(la.boxes_dst + (size_t)w * la.tiles_j + i0 / LARGE_TJ);
            // lane q merges component q of the tile box: x-min, x-max, y-min, y-max
            if (lane == 0) atomic_min_float(tb, cb.x);
            else if (lane == 1) atomic_max_float(tb + 1, cb.y);
            else if (lane == 2) atomic_min_float(tb + 2, cb.z);
            else atomic_max_float(tb + 3, cb.w);
        }
    }
}

// ---- host-side launch of n_steps large-crowd steps ---------------------------------------------------
template <typename PT, typename ST, bool SPLIT>
inline int launch_large_t(LargeScratch &s, HsfmArgs &a, int cluster, bool prune_tiles, bool reorder, cudaStream_t stream,
                          long long *launches, int *out_blocks, int *out_smem, bool chunk_kernel = false, int chunk_split = 0,
                          int *out_threads = nullptr)
{
    constexpr int NARR = SPLIT ? 6 : 4;
    const int Npad = (a.N + LARGE_TJ - 1) / LARGE_TJ * LARGE_TJ;
    const size_t n = (size_t)a.W * a.N, nv = (size_t)a.W * Npad;
    const size_t e2 = sizeof(ST) * 2;
    const bool prune = sizeof(PT) == 4 && prune_tiles;
    reorder = reorder && prune;
    cudaError_t e;
    if (!s.pos2 || s.Npad != Npad) {
        large_free(s);
        if ((e = cudaMalloc(&s.pos2, n * 16)) != cudaSuccess) return -2;
        if ((e = cudaMalloc(&s.vel2, n * 16)) != cudaSuccess) return -2;
        if ((e = cudaMalloc(&s.th2, n * e2)) != cudaSuccess) return -2;
        if (sizeof(ST) == 4) {
            if ((e = cudaMalloc(&s.thlo2, n * 8)) != cudaSuccess) return -2;
            if ((e = cudaMalloc(&s.thlo3, n * 8)) != cudaSuccess) return -2;
        }
        s.view_bytes = nv * NARR * sizeof(PT);
        if ((e = cudaMalloc(&s.view[0], s.view_bytes)) != cudaSuccess) return -2;
        if ((e = cudaMalloc(&s.view[1], s.view_bytes)) != cudaSuccess) return -2;
        if ((e = cudaMalloc((void **)&s.robot_slot, sizeof(RobotSlot) * a.W * LARGE_RESORT * 2)) != cudaSuccess) return -2;
        if ((e = cudaStreamCreateWithFlags(&s.aux, cudaStreamNonBlocking)) != cudaSuccess) return -2;
        if ((e = cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming)) != cudaSuccess) return -2;
        if ((e = cudaEventCreateWithFlags(&s.ev_join, cudaEventDisableTiming)) != cudaSuccess) return -2;
        if ((e = cudaMalloc((void **)&s.robot_regs, sizeof(RobotRegs) * a.W)) != cudaSuccess) return -2;
        for (int q = 0; q < 3; ++q)
            if ((e = cudaMalloc((void **)&s.boxes[q], sizeof(float4) * (size_t)a.W * (Npad / LARGE_TJ))) != cudaSuccess) return -2;
        for (int q = 0; q < 2; ++q)
            if ((e = cudaMalloc((void **)&s.chunk_boxes[q], sizeof(float4) * 3 * (size_t)a.W * (Npad / LARGE_CHUNK))) != cudaSuccess) return -2;   // + half boxes
        s.Npad = Npad;
    }
    int end_bit = 48;
    while ((1ll << (end_bit - 48)) < a.W) ++end_bit;
    if (reorder && !s.pos3) {
        if ((e = cudaMalloc(&s.pos3, n * 16)) != cudaSuccess) return -2;
        if ((e = cudaMalloc(&s.vel3, n * 16)) != cudaSuccess) return -2;
        if ((e = cudaMalloc(&s.th3, n * e2)) != cudaSuccess) return -2;
        for (int q = 0; q < 2; ++q) {
            if ((e = cudaMalloc((void **)&s.keys[q], n * sizeof(unsigned long long))) != cudaSuccess) return -2;
            if ((e = cudaMalloc((void **)&s.perm[q], n * sizeof(int))) != cudaSuccess) return -2;
        }
        cub::DeviceRadixSort::SortPairs(nullptr, s.sort_tmp_bytes, s.keys[0], s.keys[1], s.perm[0], s.perm[1], (int)n, 0, end_bit, stream);
        if ((e = cudaMalloc(&s.sort_tmp, s.sort_tmp_bytes ? s.sort_tmp_bytes : 1)) != cudaSuccess) return -2;
    }

    // j-split of the chunk kernel: asked for, or by the number of chunks (measured on one B200, 148 SMs: 512 chunks 15.0 / 13.2 /
    // 11.6 us per step with 1 / 2 / 4 warps per chunk, 1024 chunks 16.9 / 16.6 / 17.9, 2048 chunks 23.0 / 25.8 / 29.7)
    const long long n_chunks_all = (long long)a.W * ((a.N + LARGE_CHUNK - 1) / LARGE_CHUNK);
    const int ks = chunk_split == 1 || chunk_split == 2 || chunk_split == 4 ? chunk_split
                   : (n_chunks_all <= 4LL * s.sm_count ? 4 : (n_chunks_all <= 8LL * s.sm_count ? 2 : 1));
    const int cpb = ks == 1 ? CHUNK_TI / LARGE_CHUNK : 1;          // chunks per chunk-kernel CTA
    const int tiles_i = (a.N + LARGE_TI - 1) / LARGE_TI, tiles_j = Npad / LARGE_TJ, tiles_c = (a.N + cpb * LARGE_CHUNK - 1) / (cpb * LARGE_CHUNK);
    bool use_chunk = false;                                        // pruned FP32 regime: barrier-free warp-per-chunk kernel
    if constexpr (sizeof(ST) == 4) use_chunk = chunk_kernel && prune;
    // grid must be a multiple of the cluster size and a cluster must stay inside one world
    while (cluster > 1 && (tiles_i % cluster != 0 || LARGE_TJ % cluster != 0)) cluster >>= 1;
    const size_t smem = LARGE_STAGES * NARR * LARGE_TJ * sizeof(PT) + LARGE_STAGES * sizeof(uint64_t) + sizeof(int) * (tiles_j + 1);
    auto kern = hsfm_large_kernel<PT, ST, SPLIT>;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
    const unsigned gn = (unsigned)((n + 255) / 256), gv = (unsigned)((nv + 255) / 256);
    const float inv_strip = 1.0f / (LARGE_STRIP_SCALE * sqrtf(a.warp_cut2));

    // working state: the caller's arrays, or the spatially sorted copy (home[0] is then s.pos3 ...)
    const bool comp = sizeof(ST) == 4 && a.comp;                  // FP32: a.vel (= a.pos_lo) and a.th_lo hold the low parts
    void *pos[2] = {a.pos, s.pos2}, *vel[2] = {a.vel, s.vel2}, *th[2] = {a.th, s.th2}, *thlo[2] = {a.th_lo, s.thlo2};
    if (reorder) { pos[0] = s.pos3; vel[0] = s.vel3; th[0] = s.th3; thlo[0] = s.thlo3; }
    int cur = 0;                                                   // buffer that holds the current state
    if (!reorder) s.sorted_valid = false;
    for (int t0 = 0, t1 = 0; t0 < a.n_steps; t0 = t1) {
        // go on in the existing sorted copy while it is fresh, else sort again
        const bool cont = reorder && s.sorted_valid && s.since_sort < LARGE_RESORT;
        const int room = cont ? LARGE_RESORT - s.since_sort : LARGE_RESORT;
        t1 = t0 + room < a.n_steps ? t0 + room : a.n_steps;
        // the robots of the whole window, ahead of the pedestrians, on the side stream (slot buffers alternate per window)
        RobotSlot *const slots = s.robot_slot + (size_t)(s.window++ & 1) * LARGE_RESORT * a.W;
        if (a.robot_model != ROBOT_NONE) {
            cudaEventRecord(s.ev_fork, stream);
            cudaStreamWaitEvent(s.aux, s.ev_fork, 0);
            robot_window_kernel<<<(a.W + 127) / 128, 128, 0, s.aux>>>(a, s.robot_regs, slots, t0, t1);
            ++*launches;
            if (a.sp.robot_on) { sensor_window_kernel<<<a.W, 32, 0, s.aux>>>(a, slots, t0, t1); ++*launches; }
            cudaEventRecord(s.ev_join, s.aux);
        }
        if (cont) cur = s.cur;
        if (reorder && !cont) {
            // sort by (world, strip, y) of the CURRENT positions (in the caller's arrays), gather the state
            large_sortkey_kernel<ST><<<gn, 256, 0, stream>>>(a, inv_strip, s.keys[0], s.perm[0]);
            cub::DeviceRadixSort::SortPairs(s.sort_tmp, s.sort_tmp_bytes, s.keys[0], s.keys[1], s.perm[0], s.perm[1], (int)n, 0, end_bit, stream);
            large_permute_kernel<ST, true><<<gn, 256, 0, stream>>>(n, s.perm[1], a.pos, a.vel, a.th, pos[0], vel[0], th[0],
                                                                   comp ? a.th_lo : nullptr, thlo[0]);
            *launches += 3;
            cur = 0;
            s.since_sort = 0;
        }
        // the pair view is rebuilt from the working state (the host may have changed it between calls) -- unless the call goes
        // on in the sorted copy of the previous one, whose last step left the view (and, chunk kernel, the boxes) up to date
        const bool keep_boxes = cont && use_chunk && s.view_valid;
        if (!(cont && s.view_valid)) {
            HsfmArgs av = a;
            av.pos = pos[cur]; av.vel = vel[cur]; av.th = th[cur];
            large_view_kernel<PT, ST, SPLIT><<<gv, 256, 0, stream>>>(av, reinterpret_cast<PT *>(s.view[cur]), nv, Npad);
            if (t0 == 0) large_view_kernel<PT, ST, SPLIT><<<gv, 256, 0, stream>>>(av, reinterpret_cast<PT *>(s.view[cur ^ 1]), nv, Npad);
            *launches += 2;
        }
        const int bq0 = keep_boxes ? s.bq : 0, cq0 = keep_boxes ? s.cq : 0;
        if (a.robot_model != ROBOT_NONE) cudaStreamWaitEvent(stream, s.ev_join, 0);
        for (int t = t0; t < t1; ++t) {
            const int src = cur, dst = cur ^ 1;
            // tile boxes rotate through three buffers in the chunk-kernel regime (read, merged into, reset), chunk boxes through two
            const int bq = use_chunk ? (bq0 + t - t0) % 3 : 0, cq = use_chunk ? (cq0 + t - t0) & 1 : 0;
            LargeArgs la{};
            la.pos_src = pos[src]; la.vel_src = vel[src]; la.th_src = th[src];
            la.pos_dst = pos[dst]; la.vel_dst = vel[dst]; la.th_dst = th[dst];
            la.thlo_src = thlo[src]; la.thlo_dst = thlo[dst];
            la.view_src = s.view[src]; la.view_dst = s.view[dst];
            la.view_stride = nv; la.Npad = Npad; la.tiles_i = tiles_i; la.tiles_j = tiles_j; la.cluster = cluster;
            la.step = t; la.last = (t == a.n_steps - 1); la.robot_slot = slots + (size_t)(t - t0) * a.W;
            la.boxes = prune ? s.boxes[bq] : nullptr; la.chunk_boxes = s.chunk_boxes[cq]; la.cut2 = a.warp_cut2;
            la.boxes_dst = s.boxes[(bq + 1) % 3]; la.boxes_reset = s.boxes[(bq + 2) % 3]; la.chunk_boxes_dst = s.chunk_boxes[cq ^ 1];
            la.tiles_c = tiles_c;
            la.perm = reorder ? s.perm[1] : nullptr;
            if (prune && (!use_chunk || (t == t0 && !keep_boxes))) {   // the chunk kernel keeps the boxes up to date itself after the first step
                if constexpr (sizeof(PT) == 4)
                    large_box_kernel<PT><<<(unsigned)(a.W * tiles_j), 128, 0, stream>>>(a, Npad, tiles_j, reinterpret_cast<const PT *>(s.view[src]), nv,
                                                                                     s.boxes[bq], s.chunk_boxes[cq],
                                                                                     use_chunk ? s.chunk_boxes[cq ^ 1] : nullptr,
                                                                                     use_chunk ? s.boxes[(bq + 1) % 3] : nullptr);
                ++*launches;
            }
            if constexpr (sizeof(ST) == 4) {
                if (use_chunk) {
                    const unsigned gc = (unsigned)(a.W * tiles_c);
                    if (ks == 1 && n_chunks_all > 16LL * s.sm_count) hsfm_chunk_kernel<1, CHUNK_TI / LARGE_CHUNK, PMS_CHUNK_MINB_HI><<<gc, CHUNK_TI, 0, stream>>>(a, la);
                    else if (ks == 1) hsfm_chunk_kernel<1, CHUNK_TI / LARGE_CHUNK><<<gc, CHUNK_TI, 0, stream>>>(a, la);
                    else if (ks == 2) hsfm_chunk_kernel<2, 1><<<gc, 64, 0, stream>>>(a, la);
                    else hsfm_chunk_kernel<4, 1><<<gc, 128, 0, stream>>>(a, la);
                    if ((e = cudaGetLastError()) != cudaSuccess) return -2;
                    ++*launches;
                    cur = dst;
                    continue;
                }
            }
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)(a.W * tiles_i)); cfg.blockDim = dim3(LARGE_TI); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            if ((e = cudaLaunchKernelEx(&cfg, kern, a, la)) != cudaSuccess) return -2;
            ++*launches;
            cur = dst;
        }
        if (reorder) {             // back to the caller's order (host reads and the next sort use the caller's arrays)
            large_permute_kernel<ST, false><<<gn, 256, 0, stream>>>(n, s.perm[1], pos[cur], vel[cur], th[cur], a.pos, a.vel, a.th,
                                                                    comp ? thlo[cur] : nullptr, a.th_lo);
            ++*launches;
            s.sorted_valid = true; s.view_valid = true; s.cur = cur; s.since_sort += t1 - t0;
            s.bq = (bq0 + t1 - t0) % 3; s.cq = (cq0 + t1 - t0) & 1;
        }
    }
    // without re-ordering an odd number of steps leaves the newest state in the scratch buffers: copy it home
    if (!reorder && cur == 1) {
        cudaMemcpyAsync(a.pos, s.pos2, n * 16, cudaMemcpyDeviceToDevice, stream);
        if (sizeof(ST) == 8 || comp) cudaMemcpyAsync(a.vel, s.vel2, n * 16, cudaMemcpyDeviceToDevice, stream);
        if (comp) cudaMemcpyAsync(a.th_lo, s.thlo2, n * 8, cudaMemcpyDeviceToDevice, stream);
        cudaMemcpyAsync(a.th, s.th2, n * e2, cudaMemcpyDeviceToDevice, stream);
    }
    *out_blocks = use_chunk ? a.W * tiles_c : a.W * tiles_i;
    if (out_threads) *out_threads = use_chunk ? 32 * ks * cpb : LARGE_TI;
    *out_smem = (int)smem;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

} // namespace pms
