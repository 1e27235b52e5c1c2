// pms_api.cu -- C ABI of libpms_b200 (see include/pms.h).  Host-side handle management, state
// marshalling (fp64 host arrays <-> native structure-of-arrays in HBM), launch configuration and the
// extern "C" entry points.  All compute is in the CUDA kernels included below; there is no CPU path.
#include "../../include/pms.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "hsfm_small.cuh"
#include "hsfm_warp_api.cuh"
#include "hsfm_duo_api.cuh"
#include "hsfm_large.cuh"
#include "sensors.cuh"
#include "replay_api.cuh"
#include "orca.cuh"

using namespace pms;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(PMS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));        \
    } while (0)

template <typename T> cudaError_t dalloc(T **p, size_t n)
{
    if (n == 0) n = 1;
    cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, n * sizeof(T));
    return e;
}

// ---- marshalling kernels: host layout (W,N,3) fp64 <-> native SoA --------------------------------
// IO = element type of the caller's (W,N,3) arrays: double (the reference's dtype) or float (half the bytes)
template <typename ST, typename IO>
__global__ void pack_state_kernel(size_t n, const IO *poses, const IO *vels, HsfmArgs a, bool keep_lo)
{
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= n) return;
    using ST2 = typename Vec2<ST>::type;
    const double px = (double)poses[3 * g], py = (double)poses[3 * g + 1], pt = (double)poses[3 * g + 2];
    const double qx = vels ? (double)vels[3 * g] : 0.0, qy = vels ? (double)vels[3 * g + 1] : 0.0, qt = vels ? (double)vels[3 * g + 2] : 0.0;
    store_pv(a, g, (ST)px, (ST)py, (ST)qx, (ST)qy);
    ST2 t; t.x = (ST)pt; t.y = (ST)qt;
    reinterpret_cast<ST2 *>(a.th)[g] = t;
    if constexpr (sizeof(ST) == 4) {       // low parts: zero for a plain fp32 state (keep_lo false), else the fp32 rounding residue
        auto lo = [&](double v) { return keep_lo ? (float)(v - (double)(float)v) : 0.0f; };
        reinterpret_cast<float4 *>(a.pos_lo)[g] = make_float4(lo(px), lo(py), lo(qx), lo(qy));
        reinterpret_cast<float2 *>(a.th_lo)[g] = make_float2(lo(pt), lo(qt));
    }
}
template <typename ST, typename IO>
__global__ void unpack_state_kernel(size_t n, IO *poses, IO *vels, HsfmArgs a)
{
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= n) return;
    using ST2 = typename Vec2<ST>::type;
    ST x, y, vx, vy;
    load_pv(a, g, x, y, vx, vy);
    const ST2 t = reinterpret_cast<const ST2 *>(a.th)[g];
    double px = (double)x, py = (double)y, pt = (double)t.x, qx = (double)vx, qy = (double)vy, qt = (double)t.y;
    if constexpr (sizeof(ST) == 4) {       // hi + lo (the low parts are zero unless the compensated kernel ran)
        const float4 pl = reinterpret_cast<const float4 *>(a.pos_lo)[g];
        const float2 tl = reinterpret_cast<const float2 *>(a.th_lo)[g];
        px += (double)pl.x; py += (double)pl.y; pt += (double)tl.x; qx += (double)pl.z; qy += (double)pl.w; qt += (double)tl.y;
    }
    if (poses) { poses[3 * g] = (IO)px; poses[3 * g + 1] = (IO)py; poses[3 * g + 2] = (IO)pt; }
    if (vels) { vels[3 * g] = (IO)qx; vels[3 * g + 1] = (IO)qy; vels[3 * g + 2] = (IO)qt; }
}
template <typename IO>
__global__ void convert_vals_kernel(size_t n, const double2 *src, IO *dst)
{
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g < n) { dst[2 * g] = (IO)src[g].x; dst[2 * g + 1] = (IO)src[g].y; }
}
template <typename ST>
__global__ void convert_array_kernel(size_t n, const double *src, ST *dst)
{
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g < n) dst[g] = (ST)src[g];
}
// pms_step_tick: state (fp64, hi + lo), detector values, waypoint indices, masks, flags and the event count into the tick block
template <typename ST>
__global__ void tick_pack_kernel(size_t n, int W, int words, HsfmArgs a, double *poses, double *vels, double2 *vals, int *wp_index,
                                 uint32_t *coll, uint32_t *det, int *nonfinite, int *n_events)
{
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g < n) {
        using ST2 = typename Vec2<ST>::type;
        ST x, y, vx, vy;
        load_pv(a, g, x, y, vx, vy);
        const ST2 t = reinterpret_cast<const ST2 *>(a.th)[g];
        double px = (double)x, py = (double)y, pt = (double)t.x, qx = (double)vx, qy = (double)vy, qt = (double)t.y;
        if constexpr (sizeof(ST) == 4) {
            const float4 pl = reinterpret_cast<const float4 *>(a.pos_lo)[g];
            const float2 tl = reinterpret_cast<const float2 *>(a.th_lo)[g];
            px += (double)pl.x; py += (double)pl.y; pt += (double)tl.x; qx += (double)pl.z; qy += (double)pl.w; qt += (double)tl.y;
        }
        poses[3 * g] = px; poses[3 * g + 1] = py; poses[3 * g + 2] = pt;
        vels[3 * g] = qx; vels[3 * g + 1] = qy; vels[3 * g + 2] = qt;
        vals[g] = a.det_vals[g]; wp_index[g] = a.wp_index ? a.wp_index[g] : 0;
    }
    if (g < (size_t)W * words) { coll[g] = a.coll_mask[g]; det[g] = a.det_mask[g]; }
    if (g < (size_t)W) nonfinite[g] = a.nonfinite[g];
    if (g == 0) *n_events = a.ev_count ? *a.ev_count : 0;
}
// pms_sense_tick of a small batch (the drop-in classes: one world) in ONE kernel: the caller's poses and the robot pose are
// read from the pinned host block (mapped), stored as the handle's state (velocities zero, plain low parts -- what
// pms_set_ped_state + pms_set_robot_pose would leave), the collision list + detector evaluated like sense_kernel and the
// reading written into the output part of the same block -- no copies in the chain.  Grid (world, block of 256 pedestrians).
template <typename ST>
__global__ void sense_tick_kernel(const __grid_constant__ HsfmArgs a, const double *poses, const double *robot, double2 *vals_out,
                                  uint32_t *coll_out, uint32_t *det_out, bool keep_lo)
{
    const int w = blockIdx.x, i = blockIdx.y * blockDim.x + threadIdx.x;
    const int N = a.N, words = (N + 31) >> 5;
    const RobotSlot rb{robot[3 * w], robot[3 * w + 1], robot[3 * w + 2], 0.0, 0.0};
    if (i < 3 && blockIdx.y == 0) a.robot_pose[3 * w + i] = robot[3 * w + i];
    SenseOut so{false, false, 0.0, 0.0};
    const size_t g = (size_t)w * N + i;
    if (i < N) {
        using ST2 = typename Vec2<ST>::type;
        const double px = poses[3 * g], py = poses[3 * g + 1], pt = poses[3 * g + 2];
        const ST x = (ST)px, y = (ST)py;
        store_pv(a, g, x, y, (ST)0, (ST)0);
        ST2 t; t.x = (ST)pt; t.y = (ST)0;
        reinterpret_cast<ST2 *>(a.th)[g] = t;
        if constexpr (sizeof(ST) == 4) {       // low parts like pack_state_kernel
            auto lo = [&](double v) { return keep_lo ? (float)(v - (double)(float)v) : 0.0f; };
            const float xl = lo(px), yl = lo(py);
            reinterpret_cast<float4 *>(a.pos_lo)[g] = make_float4(xl, yl, 0.f, 0.f);
            reinterpret_cast<float2 *>(a.th_lo)[g] = make_float2(lo(pt), 0.f);
            so = sense_ped<double>(a, rb, (double)x + (double)xl, (double)y + (double)yl, true);
        } else {
            so = sense_ped<ST>(a, rb, x, y, true);
        }
        const double2 v = a.det_enabled ? make_double2(so.v0, so.v1) : a.det_vals[g];
        if (a.det_enabled) a.det_vals[g] = v;
        vals_out[g] = v;
    }
    const unsigned dm = __ballot_sync(0xffffffffu, so.det);
    const unsigned cm = __ballot_sync(0xffffffffu, so.coll);
    if ((i & 31) == 0 && (i >> 5) < words) {
        const size_t m = (size_t)w * words + (i >> 5);
        a.det_mask[m] = dm; a.coll_mask[m] = cm;
        det_out[m] = dm; coll_out[m] = cm;
    }
}
template <typename ST>
__global__ void convert_back_kernel(size_t n, const ST *src, double *dst)
{
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g < n) dst[g] = (double)src[g];
}
// current waypoint <- table[index] (FixedWaypointTracker._get_current_waypoints)
template <typename ST>
__global__ void gather_waypoints_kernel(size_t n, int n_table, const ST *table, const int *index, ST *wp)
{
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= n) return;
    wp[2 * g] = table[(g * n_table + index[g]) * 2];
    wp[2 * g + 1] = table[(g * n_table + index[g]) * 2 + 1];
}
template <typename ST>
__global__ void append_queue_kernel(int n, const int *world_ped, const double *pts, int N, int q_cap, ST *queue,
                                    const int *q_head, int *q_count, int *overflow)
{
    // sequential: several appends may target the same pedestrian and must keep their order
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int k = 0; k < n; ++k) {
        const size_t g = (size_t)world_ped[2 * k] * N + world_ped[2 * k + 1];
        if (q_count[g] >= q_cap) { *overflow = 1; continue; }
        const int slot = (q_head[g] + q_count[g]) % q_cap;
        queue[(g * q_cap + slot) * 2] = (ST)pts[2 * k];
        queue[(g * q_cap + slot) * 2 + 1] = (ST)pts[2 * k + 1];
        q_count[g]++;
    }
}
__global__ void fill_int_kernel(size_t n, int *p, int v)
{
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g < n) p[g] = v;
}

inline unsigned nblk(size_t n, unsigned t = 256) { return (unsigned)((n + t - 1) / t); }

} // namespace

// ---------------------------------------------------------------------------------------------
struct pms_sim {
    int kind = 0;          // 0 = HSFM simulation, 1 = ORCA
    int device = 0;
    int W = 0, N = 0, Np = 0, words = 0;
    int precision = PMS_FP32;
    bool compensated = false;   // PMS_FP32 (not PMS_FP32_PLAIN): hi + lo state in the warp-per-world kernel
    DetNoise det_noise{};       // device-side detector noise; det_noise.return_type = the caller's return type
    LidarNoise lidar_noise{};
    OmniNoise omni_noise{};
    uint32_t *noisy_mask = nullptr;
    double2 *noisy_vals = nullptr;
    size_t esz = 4;        // bytes of one state scalar
    cudaStream_t stream = nullptr;
    HsfmArgs a{};          // device pointers + configuration (the kernel argument block)
    // staging (device, fp64) for host <-> device marshalling
    double *stage = nullptr;
    // pms_step_tick: one device block + one pinned host block that carry every output of a tick in ONE transfer
    unsigned char *tick_dev = nullptr, *tick_host = nullptr;
    size_t tick_bytes = 0;
    size_t stage_elems = 0;
    double *controls_dev = nullptr;
    size_t controls_cap = 0;
    double *map_dev = nullptr;
    size_t map_cap = 0;
    int *overflow_dev = nullptr;
    double *noise_dev = nullptr;
    double *robot_block = nullptr;   // pose (3W) | velocity (3W) | control (2W) | rear axle (3W) doubles | world-collision (W ints)
    // snapshot
    std::vector<std::pair<void *, size_t>> snap_src;
    std::vector<void *> snap_dst;
    long long snap_steps = 0;
    bool has_snapshot = false;
    // tuning / introspection
    int tune_S = 0, tune_wpc = 0, tune_chunks = 0, tune_cluster = 0;
    bool force_large = false;
    bool no_prune = false;   // large-crowd kernel: visit every j-tile (comparison / tests)
    bool no_reorder = false; // large-crowd kernel: keep the caller's pedestrian order (no sorted working copy)
    bool async_io = false;   // state / reading transfers do not synchronise (pms_set_async_io)
    int sm_count = 148;
    int sched_cap = 0;
    pms_launch_info info{};
    bool waypoints_set = false;
    // large-crowd path scratch
    LargeScratch large{};
    // sensors inside a fused launch (pms_sensor_schedule): scheduler state, fire tables and the records of the last launch
    struct Sched {
        bool enabled[SENSOR_KINDS] = {false, false, false, false};
        double period[SENSOR_KINDS] = {0, 0, 0, 0};
        int flags[SENSOR_KINDS] = {0, 0, 0, 0};
        double hold_host[SENSOR_KINDS] = {0, 0, 0, 0};   // host mirror of hold_dev (same fp64 recurrence): sizes the records
        double snap_hold[SENSOR_KINDS] = {0, 0, 0, 0};
        int n_fires[SENSOR_KINDS] = {0, 0, 0, 0};         // readings of the last launch (host mirror)
        double *hold_dev = nullptr;                       // (SENSOR_KINDS)
        int *count_dev = nullptr;                         // (SENSOR_KINDS)
        int *tables = nullptr;                            // fire[kind][t] | steps[kind][slot], 2 * SENSOR_KINDS * table_cap ints
        int table_cap = 0;
        // records
        uint32_t *det_mask = nullptr, *coll_mask = nullptr; size_t det_mask_cap = 0, coll_mask_cap = 0;   // capacities in records
        double2 *det_vals = nullptr; size_t det_vals_cap = 0;
        int32_t *lidar_hit = nullptr; double2 *lidar_pts = nullptr; size_t lidar_cap = 0, lidar_pts_cap = 0;
        double *omni = nullptr; size_t omni_cap = 0;
        // lidar geometry
        int lidar_beams = 0, lidar_samples = 0;
        double lidar_res = 0.0, lidar_max = 0.0;
        double *lidar_dirs = nullptr;
        bool any() const { return enabled[0] || enabled[1] || enabled[2] || enabled[3]; }
    } sched;
    // recorded trajectories (pms_set_replay): frames resident in HBM, frame counter (the same for every world)
    double *replay_poses = nullptr, *replay_vels = nullptr;
    int replay_frames = 0, replay_loop = 0, replay_frame = 0, snap_replay_frame = 0;
    // ORCA
    OrcaState orca{};
};

namespace {

int ensure_stage(pms_sim *h, size_t elems)
{
    if (elems <= h->stage_elems) return 0;
    if (h->stage) cudaFree(h->stage);
    h->stage = nullptr;
    h->stage_elems = 0;
    CK(cudaMalloc((void **)&h->stage, elems * sizeof(double)));
    h->stage_elems = elems;
    return 0;
}

struct Guard {
    int prev = -1;
    explicit Guard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename F> int by_state(pms_sim *h, F &&f)
{
    if (h->precision == PMS_FP32) return f(float{});
    return f(double{});
}

// state of all pedestrians from / to (W,N,3) DEVICE arrays of doubles or floats, on the handle's stream
int state_from_device(pms_sim *h, const void *poses, const void *vels, bool f32)
{
    const size_t n = (size_t)h->W * h->N;
    const unsigned nb_ = (unsigned)((n + 255) / 256);
    if (h->precision == PMS_FP32) {
        if (f32) pack_state_kernel<float, float><<<nb_, 256, 0, h->stream>>>(n, (const float *)poses, (const float *)vels, h->a, h->compensated);
        else pack_state_kernel<float, double><<<nb_, 256, 0, h->stream>>>(n, (const double *)poses, (const double *)vels, h->a, h->compensated);
    } else {
        if (f32) pack_state_kernel<double, float><<<nb_, 256, 0, h->stream>>>(n, (const float *)poses, (const float *)vels, h->a, false);
        else pack_state_kernel<double, double><<<nb_, 256, 0, h->stream>>>(n, (const double *)poses, (const double *)vels, h->a, false);
    }
    CK(cudaGetLastError());
    h->large.sorted_valid = false;            // the large-crowd kernel's sorted working copy no longer mirrors the state
    return 0;
}
int state_to_device(pms_sim *h, void *poses, void *vels, bool f32)
{
    const size_t n = (size_t)h->W * h->N;
    const unsigned nb_ = (unsigned)((n + 255) / 256);
    if (h->precision == PMS_FP32) {
        if (f32) unpack_state_kernel<float, float><<<nb_, 256, 0, h->stream>>>(n, (float *)poses, (float *)vels, h->a);
        else unpack_state_kernel<float, double><<<nb_, 256, 0, h->stream>>>(n, (double *)poses, (double *)vels, h->a);
    } else {
        if (f32) unpack_state_kernel<double, float><<<nb_, 256, 0, h->stream>>>(n, (float *)poses, (float *)vels, h->a);
        else unpack_state_kernel<double, double><<<nb_, 256, 0, h->stream>>>(n, (double *)poses, (double *)vels, h->a);
    }
    CK(cudaGetLastError());
    return 0;
}

// ---- launch of the small-world fused kernel ----
template <typename PT, typename ST, bool SPLIT, int S>
int launch_small_t(pms_sim *h, int wpc, bool big)
{
    const size_t smem = (SmallSmem<PT, ST, SPLIT>::bytes(h->Np, S) + (sizeof(ST) == 4 && h->a.comp ? 6 * sizeof(float) * h->Np : 0)) * wpc;
    const int threads = wpc * S * h->Np + 32;
    const int groups = (h->W + wpc - 1) / wpc;
    int blocks = groups;
    h->a.wpc = wpc;
    const bool sens = h->a.sp.on != 0;      // sensor plan: its own instantiation
    auto kern = sens ? (big ? hsfm_small_kernel_big<PT, ST, SPLIT, S, true> : hsfm_small_kernel<PT, ST, SPLIT, S, true>)
                     : (big ? hsfm_small_kernel_big<PT, ST, SPLIT, S> : hsfm_small_kernel<PT, ST, SPLIT, S>);
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // dynamic (chunk, group) scheduling when a static grid would leave a partial last wave
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    const int slots = occ * h->sm_count;
    h->a.n_chunks = 1; h->a.chunk_steps = h->a.n_steps;
    int want = h->tune_chunks;
    if (want == 0 && groups > slots && h->a.n_steps >= 64) want = h->a.n_steps / 32 < 8 ? h->a.n_steps / 32 : 8;
    if (want > 1 && slots > 0) {
        h->a.chunk_steps = (h->a.n_steps + want - 1) / want;
        h->a.n_chunks = (h->a.n_steps + h->a.chunk_steps - 1) / h->a.chunk_steps;
    }
    if (h->a.n_chunks > 1) {
        if (groups > h->sched_cap) {
            if (h->a.group_done) cudaFree(h->a.group_done);
            h->a.group_done = nullptr; h->sched_cap = 0;
            CK(cudaMalloc((void **)&h->a.group_done, sizeof(int) * (size_t)groups));
            h->sched_cap = groups;
        }
        CK(cudaMemsetAsync(h->a.group_done, 0, sizeof(int) * (size_t)groups, h->stream));
        CK(cudaMemsetAsync(h->a.work_counter, 0, sizeof(int), h->stream));
        const long long total = (long long)groups * h->a.n_chunks;
        blocks = (int)(total < slots ? total : slots);
    }
    kern<<<blocks, threads, smem, h->stream>>>(h->a);
    CK(cudaGetLastError());
    h->info.kernel = 0;
    h->info.threads_per_block = threads;
    h->info.blocks = blocks;
    h->info.worlds_per_block = wpc;
    h->info.j_split = S;
    h->info.smem_bytes = (int)smem;
    h->info.n_chunks = h->a.n_chunks;
    h->info.blocks_per_sm = occ;
    h->info.kernel_launches++;
    return 0;
}

template <int S> int launch_small_s(pms_sim *h, int wpc, bool big)
{
    switch (h->precision) {
    case PMS_FP32: return launch_small_t<float, float, false, S>(h, wpc, big);
    case PMS_MIXED: return launch_small_t<float, double, true, S>(h, wpc, big);
    default: return launch_small_t<double, double, false, S>(h, wpc, big);
    }
}

int launch_small(pms_sim *h)
{
    const int Np = h->Np;
    int S = h->tune_S < 0 ? -h->tune_S - 1 : h->tune_S;   // j_split = -1: automatic, -1-S: forced S
    if (S <= 0) {
        // enough threads to fill the SMs; every slice keeps >= 8 pairs for ILP
        const long long target = (long long)h->sm_count * 1536;
        S = 1;
        while (S < 8 && (long long)h->W * Np * S * 2 <= target && h->N / (S * 2) >= 8) S *= 2;
    }
    while (S > 1 && S * Np + 32 > 1024) S /= 2;
    if (S != 1 && S != 2 && S != 4 && S != 8) return fail(PMS_ERR_INVALID, "j_split must be 1, 2, 4 or 8");
    int wpc = h->tune_wpc;
    if (wpc <= 0) wpc = 256 / (S * Np);
    if (wpc < 1) wpc = 1;
    const int wpc_max = (S * Np == 32) ? 31 : 15;   // named barriers 1..15 (single-warp worlds need none)
    if (wpc > wpc_max) wpc = wpc_max;
    if (wpc > h->W) wpc = h->W;
    while (wpc > 1 && wpc * S * Np + 32 > 1024) --wpc;
    const bool big = wpc * S * Np + 32 > 288;
    switch (S) {
    case 1: return launch_small_s<1>(h, wpc, big);
    case 2: return launch_small_s<2>(h, wpc, big);
    case 4: return launch_small_s<4>(h, wpc, big);
    default: return launch_small_s<8>(h, wpc, big);
    }
}

// ---- launch of the warp-per-world symmetric-pair kernel (hsfm_warp.cu): fp32 pair modes, N <= 128 ----
int launch_warp(pms_sim *h)
{
    int wpc = h->tune_wpc > 0 ? h->tune_wpc : 4;
    if (wpc > WARP_MAX_WPC) wpc = WARP_MAX_WPC;
    if (wpc > h->W) wpc = h->W;
    // register budget of the instantiation (WarpCfg in hsfm_warp.cuh): j_split 0 = 128 registers, 1 = 96, 2 = 80, 3 = 80 (256 threads)
    const bool sens = h->a.sp.on != 0;      // sensor plan: its own instantiation (register budget 0)
    const int T = h->Np / 32, minb = !sens && h->tune_S >= 1 && h->tune_S <= 3 ? h->tune_S : 0;   // 4..7 select / exclude hsfm_duo
    if ((minb == 1 || minb == 2) && wpc > 4) wpc = 4;
    const bool split = h->precision == PMS_MIXED;
    const bool comp = h->precision == PMS_FP32 && h->compensated;
    const int threads = wpc * 32;
    const int groups = (h->W + wpc - 1) / wpc;
    int blocks = groups, occ = 0;
    size_t smem = 0;
    h->a.wpc = wpc;
    CK(warp_kernel_prepare(split, comp, sens, T, minb, wpc, &occ, &smem));
    const int slots = occ * h->sm_count;
    h->a.n_chunks = 1; h->a.chunk_steps = h->a.n_steps;
    int want = h->tune_chunks;
    if (want == 0 && groups > slots && h->a.n_steps >= 64) want = h->a.n_steps / 32 < 8 ? h->a.n_steps / 32 : 8;
    if (want > 1 && slots > 0) {
        h->a.chunk_steps = (h->a.n_steps + want - 1) / want;
        h->a.n_chunks = (h->a.n_steps + h->a.chunk_steps - 1) / h->a.chunk_steps;
    }
    if (h->a.n_chunks > 1) {
        if (groups > h->sched_cap) {
            if (h->a.group_done) cudaFree(h->a.group_done);
            h->a.group_done = nullptr; h->sched_cap = 0;
            CK(cudaMalloc((void **)&h->a.group_done, sizeof(int) * (size_t)groups));
            h->sched_cap = groups;
        }
        CK(cudaMemsetAsync(h->a.group_done, 0, sizeof(int) * (size_t)groups, h->stream));
        CK(cudaMemsetAsync(h->a.work_counter, 0, sizeof(int), h->stream));
        const long long total = (long long)groups * h->a.n_chunks;
        blocks = (int)(total < slots ? total : slots);
    }
    CK(warp_kernel_launch(split, comp, sens, T, minb, h->a, blocks, threads, smem, h->stream));
    if (h->precision == PMS_FP32 && !comp) {        // plain fp32 state: the low parts a set_state left are gone
        const size_t n = (size_t)h->W * h->N;
        CK(cudaMemsetAsync(h->a.pos_lo, 0, n * 16, h->stream));
        CK(cudaMemsetAsync(h->a.th_lo, 0, n * 8, h->stream));
    }
    h->info.kernel = 3;
    h->info.threads_per_block = threads;
    h->info.blocks = blocks;
    h->info.worlds_per_block = wpc;
    h->info.j_split = 0;
    h->info.smem_bytes = (int)smem;
    h->info.n_chunks = h->a.n_chunks;
    h->info.blocks_per_sm = occ;
    h->info.kernel_launches++;
    return 0;
}

// ---- launch of the pair-warp / owner-warp kernel (hsfm_duo.cu): FP32 modes, N <= 64, batches that leave the schedulers
// under-filled with one warp per world ----
bool duo_eligible(const pms_sim *h)
{
    return h->precision == PMS_FP32 && h->Np <= 32 * DUO_MAX_TILES;
}

// automatic choice: the warp-per-world kernel brings ONE warp per world to the schedulers; below ~2 warps per scheduler its
// dependent chain per step is exposed and the pair-warp / owner-warp kernel (3T + 1 warps per world, one world per CTA,
// 10 / 4 - 7 CTAs resident per SM for N <= 32 / 64) wins.  Measured cross-over on one B200 (profiles/r2_duo.md): N <= 32: one
// wave of CTAs (1024 worlds: 1.04 vs 1.46 ms per 1000 steps, 2048: 1.97 vs 1.97); N <= 64: up to ~2.5 waves (512: 1.48 vs
// 2.81, 1024: 2.52 vs 3.27; 2048 worlds would take 4 waves and lose to the 4.8 ms of the warp-per-world kernel).
// Two-tile worlds: which variant (one or two units of the pair scheme per pair warp: 7 warps and 4 worlds per SM, or 5 warps
// and 5 worlds per SM) and whether the kernel beats the warp-per-world one is a matter of how the batch fills the SMs wave by
// wave.  Model from measurements on one B200 (profiles/r2_history.md, tools/ab_duo_variants.py; ms per 1000 steps): an SM that
// holds n worlds takes t[n]; the fullest SM holds ceil(W / SMs) worlds, in waves of the variant's occupancy; consecutive waves
// overlap by about 7 %.  The warp-per-world kernel's time moves in layers of one more CTA (4 worlds) per SM.  Checked against 20
// measured batch sizes from 128 to 2960 worlds: the choice is the fastest of the three everywhere but at 1184 (3 % off).
double duo_model_ms(int n, int variant)
{
    static const double t[DUO_VARIANTS + 1][8] = {{0.0},
        {0.0, 0.755, 0.956, 1.211, 1.460},                          // 1: one unit per pair warp, 4 worlds per SM
        {0.0, 0.897, 1.093, 1.239, 1.396, 1.764},                   // 2: two units, 5 per SM
        {0.0, 1.039, 1.213, 1.353, 1.468, 1.864, 2.167, 2.282},     // 3: two units, 7 per SM
        {0.0, 0.783, 1.001, 1.161, 1.319},                          // 4: two units, 4 per SM
        {0.0, 0.734, 0.960, 1.125}};                                // 5: two units, 3 per SM
    const int occ = duo_variant_blocks(variant);
    double s = 0.0;
    int waves = 0;
    for (; n > 0; ++waves) {
        const int k = n < occ ? n : occ;
        s += t[variant][k];
        n -= k;
    }
    return waves > 1 ? (variant == 3 ? 0.965 : 0.93) * s : s;
}
double warp_model_ms(int W, int sms)
{
    static const double layer_ms[5] = {0.0, 2.46, 3.03, 3.83, 4.75};
    const int layers = (W + 4 * sms - 1) / (4 * sms);       // one more CTA (4 worlds) on every SM
    return layers <= 4 ? layer_ms[layers] : 1.13 * layers;
}

// variant of the kernel (hsfm_duo_api.cuh): one tile: 2; two tiles: the fastest by the model above (forced by j_split 4 +
// worlds_per_block 1 / 2 / 3: tests, A/B runs)
int duo_best_variant(int n)
{
    int best = 1;
    for (int v = 2; v <= DUO_VARIANTS; ++v)
        if (duo_model_ms(n, v) < duo_model_ms(n, best)) best = v;
    return best;
}
int duo_upw(const pms_sim *h)
{
    if (h->Np <= 32) return 2;
    if (h->tune_S == 4 && h->tune_wpc >= 1 && h->tune_wpc <= DUO_VARIANTS) return h->tune_wpc;
    return duo_best_variant((h->W + h->sm_count - 1) / h->sm_count);
}

bool duo_preferred(const pms_sim *h)
{
    if (h->tune_S == 5) return false;                      // forced: warp-per-world kernel
    if (h->tune_S == 4) return true;                       // forced: pair-warp / owner-warp kernel
    if (h->tune_S != 0 || h->tune_wpc > 0 || h->tune_chunks > 0) return false;   // explicit launch knobs of the other kernels
    int occ = h->Np <= 32 ? 10 : 4;                        // resident CTAs (= worlds) per SM of the instantiation
    size_t smem = 0;
    if (duo_kernel_prepare(h->compensated, false, h->Np / 32, h->Np <= 32 ? 2 : 1, &occ, &smem) != cudaSuccess || occ < 1) return false;
    const long long wave = (long long)occ * h->sm_count;
    // measured cross-over against the warp-per-world kernel (profiles/r2_history.md): one tile: 1.65 waves of 10 worlds per SM
    // (2048 x 32: 1.80 vs 1.97 ms, 2664 x 32: 2.23 vs 2.18)
    if (h->Np <= 32) return 20LL * h->W <= 33 * wave;
    if (occ < 4) return 2LL * h->W <= 5 * wave;            // not the occupancy the model was measured at
    const int n = (h->W + h->sm_count - 1) / h->sm_count;
    return duo_model_ms(n, duo_best_variant(n)) < warp_model_ms(h->W, h->sm_count);
}

int launch_duo(pms_sim *h)
{
    const bool sens = h->a.sp.on != 0;      // sensor plan: its own instantiation (no other-occupancy builds of it: variants 3 .. 5 run as 2)
    const int T = h->Np / 32, upw = sens && duo_upw(h) >= 3 ? 2 : duo_upw(h);
    const bool comp = h->compensated;
    int occ = 0;
    size_t smem = 0;
    h->a.wpc = 1;
    h->a.n_chunks = 1; h->a.chunk_steps = h->a.n_steps;
    CK(duo_kernel_prepare(comp, sens, T, upw, &occ, &smem));
    CK(duo_kernel_launch(comp, sens, T, upw, h->a, h->W, smem, h->stream));
    if (!comp) {        // plain fp32 state: the low parts a set_state left are gone
        const size_t n = (size_t)h->W * h->N;
        CK(cudaMemsetAsync(h->a.pos_lo, 0, n * 16, h->stream));
        CK(cudaMemsetAsync(h->a.th_lo, 0, n * 8, h->stream));
    }
    h->info.kernel = 4;
    h->info.threads_per_block = duo_threads(T, duo_variant_upw(upw));
    h->info.blocks = h->W;
    h->info.worlds_per_block = 1;
    h->info.j_split = upw;
    h->info.smem_bytes = (int)smem;
    h->info.n_chunks = 1;
    h->info.blocks_per_sm = occ;
    h->info.kernel_launches++;
    return 0;
}

// large crowds: one launch per step, cluster-multicast j-tiles (hsfm_large.cuh)
int launch_large(pms_sim *h)
{
    int cluster = h->tune_cluster > 0 ? h->tune_cluster : 4;
    if (cluster != 1 && cluster != 2 && cluster != 4 && cluster != 8) return fail(PMS_ERR_INVALID, "cluster size must be 1, 2, 4 or 8");
    int blocks = 0, smem = 0, threads = LARGE_TI, rc;
    long long launches = 0;
    switch (h->precision) {
    // FP32 with exact pruning runs the barrier-free chunk kernel unless a cluster size was asked for explicitly
    case PMS_FP32: rc = launch_large_t<float, float, false>(h->large, h->a, cluster, !h->no_prune, !h->no_reorder, h->stream, &launches, &blocks, &smem,
                                                            h->tune_cluster == 0, h->tune_S, &threads); break;
    case PMS_MIXED: rc = launch_large_t<float, double, true>(h->large, h->a, cluster, !h->no_prune, !h->no_reorder, h->stream, &launches, &blocks, &smem); break;
    default: rc = launch_large_t<double, double, false>(h->large, h->a, cluster, false, false, h->stream, &launches, &blocks, &smem); break;
    }
    if (rc) return fail(PMS_ERR_CUDA, std::string("large-crowd launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    h->info.kernel = 1;
    h->info.threads_per_block = threads;
    h->info.blocks = blocks;
    h->info.worlds_per_block = 0;
    h->info.j_split = cluster;
    h->info.smem_bytes = smem;
    h->info.n_chunks = 1;
    h->info.blocks_per_sm = 0;
    h->info.kernel_launches += launches;
    return 0;
}

// np.linspace(0, 2*pi, n_beams) -> cos, sin on the host with the C library (sensors/_lidar.py:111,113)
void lidar_directions(int n_beams, std::vector<double> &dirs)
{
    dirs.resize((size_t)n_beams * 2);
    for (int k = 0; k < n_beams; ++k) {
        double ang;
        if (n_beams == 1) ang = 0.0;
        else if (k == n_beams - 1) ang = 2.0 * PMS_PI;
        else ang = (double)k * ((2.0 * PMS_PI - 0.0) / (double)(n_beams - 1)) + 0.0;
        dirs[2 * k] = std::cos(ang); dirs[2 * k + 1] = std::sin(ang);
    }
}

template <typename T> int grow(T **p, size_t *cap, size_t need_records, size_t per_record)
{
    if (need_records <= *cap) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    CK(cudaMalloc((void **)p, (need_records * per_record ? need_records * per_record : 1) * sizeof(T)));
    *cap = need_records;
    return 0;
}

// Sensors inside the launch that follows (pms_sensor_schedule): replay the hold-time scheduler for its n_steps on the device
// (sched_kernel) and on the host (the mirror sizes the records), and fill the SensorPlan of the kernel argument block.
int prepare_sensor_plan(pms_sim *h, int n_steps, bool zero_masks)
{
    pms_sim::Sched &sc = h->sched;
    HsfmArgs &a = h->a;
    a.sp = SensorPlan{};
    if (!sc.any()) return 0;
    if (sc.enabled[SENSOR_LIDAR] && sc.lidar_beams <= 0) return fail(PMS_ERR_STATE, "lidar is scheduled but pms_lidar_config was not called");
    if (!sc.hold_dev) {
        CK(dalloc(&sc.hold_dev, (size_t)SENSOR_KINDS));
        CK(dalloc(&sc.count_dev, (size_t)SENSOR_KINDS));
        CK(cudaMemcpyAsync(sc.hold_dev, sc.hold_host, sizeof(double) * SENSOR_KINDS, cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));      // hold_host is a member, but keep the source stable for the copy anyway
    }
    if (n_steps > sc.table_cap) {
        if (sc.tables) cudaFree(sc.tables);
        sc.tables = nullptr; sc.table_cap = 0;
        const int cap = (n_steps + 255) & ~255;
        CK(cudaMalloc((void **)&sc.tables, sizeof(int) * 2 * SENSOR_KINDS * (size_t)cap));
        sc.table_cap = cap;
    }
    // host mirror of the recurrence (core/_simulation.py:177-180): IEEE fp64 additions, identical on both sides
    double new_hold[SENSOR_KINDS];
    int new_fires[SENSOR_KINDS];
    SchedArgs sa{};
    sa.n_steps = n_steps; sa.dt = a.dt; sa.hold = sc.hold_dev; sa.count = sc.count_dev;
    for (int k = 0; k < SENSOR_KINDS; ++k) {
        sa.enabled[k] = sc.enabled[k] ? 1 : 0; sa.period[k] = sc.period[k];
        sa.fire[k] = sc.tables + (size_t)k * sc.table_cap;
        sa.steps[k] = sc.tables + (size_t)(SENSOR_KINDS + k) * sc.table_cap;
        new_hold[k] = sc.hold_host[k]; new_fires[k] = 0;
        if (!sc.enabled[k]) continue;
        volatile double hold = sc.hold_host[k];
        int n = 0;
        for (int t = 0; t < n_steps; ++t) {
            hold = hold + a.dt;
            if (hold >= sc.period[k]) { ++n; hold = 0.0; }
        }
        new_hold[k] = hold; new_fires[k] = n;
    }
    // the mirror is committed only once the device scheduler has been launched: a failed allocation below leaves both sides as they were
    auto commit = [&]() { for (int k = 0; k < SENSOR_KINDS; ++k) { sc.hold_host[k] = new_hold[k]; sc.n_fires[k] = new_fires[k]; } };
    const int *nf = new_fires;
    const size_t W = h->W, N = h->N, words = h->words;
    if (sc.enabled[SENSOR_DETECTOR]) {
        if (grow(&sc.det_mask, &sc.det_mask_cap, (size_t)nf[SENSOR_DETECTOR], W * words)) return PMS_ERR_CUDA;
        if ((sc.flags[SENSOR_DETECTOR] & PMS_SENSOR_VALUES) && grow(&sc.det_vals, &sc.det_vals_cap, (size_t)nf[SENSOR_DETECTOR], W * N)) return PMS_ERR_CUDA;
    }
    if (sc.enabled[SENSOR_COLLISIONS] && grow(&sc.coll_mask, &sc.coll_mask_cap, (size_t)nf[SENSOR_COLLISIONS], W * words)) return PMS_ERR_CUDA;
    if (sc.enabled[SENSOR_LIDAR]) {
        if (grow(&sc.lidar_hit, &sc.lidar_cap, (size_t)nf[SENSOR_LIDAR], W * sc.lidar_beams)) return PMS_ERR_CUDA;
        if (grow(&sc.lidar_pts, &sc.lidar_pts_cap, (size_t)nf[SENSOR_LIDAR], W * sc.lidar_beams)) return PMS_ERR_CUDA;
    }
    if (sc.enabled[SENSOR_OMNI] && grow(&sc.omni, &sc.omni_cap, (size_t)nf[SENSOR_OMNI], W)) return PMS_ERR_CUDA;
    sched_kernel<<<1, 32, 0, h->stream>>>(sa);
    CK(cudaGetLastError());
    commit();
    h->info.kernel_launches++;
    SensorPlan &sp = a.sp;
    const bool has_robot = a.robot_model != ROBOT_NONE;
    sp.on = 1;
    sp.ped_on = (sc.enabled[SENSOR_DETECTOR] || sc.enabled[SENSOR_COLLISIONS]) ? 1 : 0;
    sp.robot_on = ((sc.enabled[SENSOR_LIDAR] || sc.enabled[SENSOR_OMNI]) && has_robot) ? 1 : 0;
    for (int k = 0; k < SENSOR_KINDS; ++k) sp.fire[k] = sc.enabled[k] ? sa.fire[k] : nullptr;
    sp.det_mask = sc.det_mask; sp.det_vals = (sc.flags[SENSOR_DETECTOR] & PMS_SENSOR_VALUES) ? sc.det_vals : nullptr;
    sp.coll_mask = sc.coll_mask;
    sp.lidar_beams = sc.lidar_beams; sp.lidar_samples = sc.lidar_samples; sp.lidar_res = sc.lidar_res; sp.lidar_dirs = sc.lidar_dirs;
    sp.lidar_hit = sc.lidar_hit; sp.lidar_pts = sc.lidar_pts; sp.lidar_noise = h->lidar_noise;
    sp.omni_dist = sc.omni; sp.omni_noise = h->omni_noise;
    // no robot: the step kernels skip the sensor pass -- the records read as empty; sorted large-crowd copy: bits are OR-ed in
    if (!has_robot || zero_masks) {
        if (sc.enabled[SENSOR_DETECTOR] && sc.n_fires[SENSOR_DETECTOR])
            CK(cudaMemsetAsync(sc.det_mask, 0, sizeof(uint32_t) * sc.n_fires[SENSOR_DETECTOR] * W * words, h->stream));
        if (sc.enabled[SENSOR_COLLISIONS] && sc.n_fires[SENSOR_COLLISIONS])
            CK(cudaMemsetAsync(sc.coll_mask, 0, sizeof(uint32_t) * sc.n_fires[SENSOR_COLLISIONS] * W * words, h->stream));
        if (!has_robot && sp.det_vals && sc.n_fires[SENSOR_DETECTOR])
            CK(cudaMemsetAsync(sc.det_vals, 0, sizeof(double2) * sc.n_fires[SENSOR_DETECTOR] * W * N, h->stream));
    }
    return 0;
}

int clear_sensor_outputs(pms_sim *h)
{
    CK(cudaMemsetAsync(h->a.det_mask, 0, sizeof(uint32_t) * (size_t)h->W * h->words, h->stream));
    CK(cudaMemsetAsync(h->a.coll_mask, 0, sizeof(uint32_t) * (size_t)h->W * h->words, h->stream));
    return 0;
}

} // namespace

// =================================================================================================
extern "C" {

const char *pms_last_error(void) { return g_err.c_str(); }
int pms_version(void) { return 100; }

void pms_default_hsfm_params(pms_hsfm_params *p)
{
    if (!p) return;
    p->tau = 0.2; p->A = 2000.0; p->B = 0.08; p->k_1 = 1.2e5; p->k_2 = 2.4e5; p->k_o = 1.0; p->k_d = 500.0;
    p->k_lambda = 0.3; p->alpha = 3.0; p->pedestrian_mass = 70.0; p->pedestrian_radius = 0.3; p->robot_radius = 0.35;
}

void pms_default_orca_params(pms_orca_params *p)
{
    if (!p) return;
    p->time_step = 0.01; p->neighbor_dist = 1.0; p->max_neighbors = 10; p->time_horizon = 2.0; p->radius = 0.3;
    p->goal_reaching_threshold = 0.2;
}

int pms_create(int device, int n_worlds, int n_peds, int precision, const pms_hsfm_params *params, pms_sim **out)
{
    if (!out) return fail(PMS_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n_worlds < 1 || n_peds < 1) return fail(PMS_ERR_INVALID, "n_worlds and n_peds must be >= 1");
    if (precision < PMS_FP32 || precision > PMS_FP32_PLAIN) return fail(PMS_ERR_INVALID, "unknown precision");
    const bool compensated = precision == PMS_FP32;
    if (precision == PMS_FP32_PLAIN) precision = PMS_FP32;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(PMS_ERR_CUDA, std::string("no CUDA device available (libpms_b200 has no CPU fallback): ") +
                                      cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(PMS_ERR_INVALID, "device index out of range");
    Guard gd(device);
    pms_hsfm_params p;
    if (params) p = *params; else pms_default_hsfm_params(&p);
    pms_sim *h = new pms_sim();
    h->device = device;
    h->W = n_worlds; h->N = n_peds; h->Np = (n_peds + 31) & ~31; h->words = (n_peds + 31) / 32;
    h->precision = precision;
    h->compensated = compensated;
    h->esz = precision == PMS_FP32 ? 4 : 8;
    const size_t n = (size_t)n_worlds * n_peds;
    HsfmArgs &a = h->a;
    a.W = n_worlds; a.N = n_peds; a.Np = h->Np; a.wpc = 1;
    a.tau = p.tau; a.A = p.A; a.B = p.B; a.k1 = p.k_1; a.k2 = p.k_2; a.ko = p.k_o; a.kd = p.k_d;
    a.k_lambda = p.k_lambda; a.alpha = p.alpha; a.mass = p.pedestrian_mass; a.ped_radius = p.pedestrian_radius;
    a.hsfm_robot_radius = p.robot_radius;
    a.robot_radius = 0.35; a.wheel_base = 1.0; a.robot_visible = 1; a.robot_model = ROBOT_NONE;
    a.reach = 0.3; a.tracker_kind = TRACKER_FIXED; a.n_table = 1;
    a.det_enabled = 0; a.det_max_dist = 3.0; a.det_half_fov = 0.5 * (30.0 * PMS_PI / 180.0); a.det_type = DET_POLAR;
#define CKC(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            pms_destroy(h);                                                                        \
            return fail(PMS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));        \
        }                                                                                          \
    } while (0)
    CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    const size_t e2 = h->esz * 2;
    unsigned char *raw;
    CKC(dalloc(&raw, n * (precision == PMS_FP32 ? 16 : 16))); a.pos = raw;
    CKC(dalloc(&raw, n * 16)); a.vel = raw;
    CKC(dalloc(&raw, n * e2)); a.th = raw;
    if (precision == PMS_FP32) {        // low parts of the state: {x,y,vx,vy} in the (otherwise unused) vel array, {theta, omega}
        a.pos_lo = a.vel;
        CKC(dalloc(&raw, n * 8)); a.th_lo = raw;
        a.comp = compensated ? 1 : 0;
    }
    CKC(dalloc(&raw, n * e2)); a.wp = raw;
    CKC(dalloc(&raw, n * h->esz)); a.vmag = raw;
    CKC(dalloc(&a.wp_index, n));
    CKC(dalloc(&a.q_head, n));
    CKC(dalloc(&a.q_count, n));
    a.ev_cap = 1 << 16;
    CKC(dalloc(&a.events, (size_t)a.ev_cap * 3));
    CKC(dalloc(&a.ev_count, 1));
    // robot pose | velocity | control | rear axle | world-collision flag in ONE allocation: pms_step_tick uploads them in one copy
    CKC(dalloc(&h->robot_block, (size_t)n_worlds * 12));
    a.robot_pose = h->robot_block; a.robot_vel = a.robot_pose + 3 * (size_t)n_worlds; a.robot_ctrl = a.robot_vel + 3 * (size_t)n_worlds;
    a.robot_rear = a.robot_ctrl + 2 * (size_t)n_worlds;
    a.robot_world_collision = reinterpret_cast<int *>(a.robot_rear + 3 * (size_t)n_worlds);
    CKC(dalloc(&a.det_mask, (size_t)n_worlds * h->words));
    CKC(dalloc(&a.coll_mask, (size_t)n_worlds * h->words));
    CKC(dalloc(&a.det_vals, n));
    CKC(dalloc(&a.det_count, (size_t)n_worlds));
    CKC(dalloc(&a.coll_count, (size_t)n_worlds));
    CKC(dalloc(&a.nonfinite, (size_t)n_worlds));
    CKC(dalloc(&h->overflow_dev, 1));
    CKC(dalloc(&a.work_counter, 1));
    {
        cudaDeviceProp prop;
        CKC(cudaGetDeviceProperties(&prop, device));
        h->sm_count = prop.multiProcessorCount;
        h->large.sm_count = h->sm_count;
    }
    fill_int_kernel<<<nblk(n_worlds), 256, 0, h->stream>>>((size_t)n_worlds, a.nonfinite, 0x7fffffff);
    // default speed 1.5 (_hsfm.py:205) and a waypoint table of one row (set properly by the caller)
    {
        std::vector<double> v(n, 1.5);
        if (ensure_stage(h, n * 6)) { pms_destroy(h); return PMS_ERR_CUDA; }
        CKC(cudaMemcpyAsync(h->stage, v.data(), n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        if (precision == PMS_FP32) convert_array_kernel<float><<<nblk(n), 256, 0, h->stream>>>(n, h->stage, (float *)a.vmag);
        else convert_array_kernel<double><<<nblk(n), 256, 0, h->stream>>>(n, h->stage, (double *)a.vmag);
        CKC(cudaStreamSynchronize(h->stream));
    }
#undef CKC
    *out = h;
    return PMS_OK;
}

int pms_destroy(pms_sim *h)
{
    if (!h) return PMS_OK;
    Guard gd(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    HsfmArgs &a = h->a;
    void *ptrs[] = {a.th_lo, a.pos, a.vel, a.th, a.wp, a.vmag, a.wp_index, (void *)a.table, (void *)a.queue, a.q_head, a.q_count,
                    a.events, a.ev_count, h->robot_block, a.det_mask, a.coll_mask, a.det_vals, a.det_count, a.coll_count,
                    a.nonfinite, h->noisy_mask, h->noisy_vals, h->stage, h->controls_dev, h->map_dev, h->overflow_dev, a.work_counter, a.group_done, h->noise_dev};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (void *p : h->snap_dst) if (p) cudaFree(p);
    if (h->tick_dev) cudaFree(h->tick_dev);
    if (h->tick_host) cudaFreeHost(h->tick_host);
    if (h->replay_poses) cudaFree(h->replay_poses);
    if (h->replay_vels) cudaFree(h->replay_vels);
    {
        pms_sim::Sched &sc = h->sched;
        void *sp_[] = {sc.hold_dev, sc.count_dev, sc.tables, sc.det_mask, sc.coll_mask, sc.det_vals, sc.lidar_hit, sc.lidar_pts, sc.omni, sc.lidar_dirs};
        for (void *p : sp_) if (p) cudaFree(p);
    }
    large_free(h->large);
    orca_free(h->orca);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return PMS_OK;
}

int pms_n_worlds(const pms_sim *h) { return h ? h->W : 0; }
int pms_n_peds(const pms_sim *h) { return h ? h->N : 0; }

#define HCHECK(h)                                                     \
    if (!(h)) return fail(PMS_ERR_INVALID, "handle is NULL");         \
    if ((h)->kind != 0) return fail(PMS_ERR_STATE, "not an HSFM handle"); \
    Guard gd__((h)->device)

// end of a transfer call: blocking unless the handle is in asynchronous-I/O mode
#define IO_DONE(h)                                                   \
    do {                                                             \
        if (!(h)->async_io) CK(cudaStreamSynchronize((h)->stream));  \
    } while (0)

int pms_set_async_io(pms_sim *h, int enabled)
{
    if (!h) return fail(PMS_ERR_INVALID, "handle is NULL");
    h->async_io = enabled != 0;
    return PMS_OK;
}

// host (W,N,3) arrays -> staging -> native SoA; esz = bytes of one host element (8: double, 4: float)
static int set_ped_state_host(pms_sim *h, const void *poses, const void *vels, size_t esz)
{
    if (!poses) return fail(PMS_ERR_INVALID, "poses is NULL");
    const size_t n = (size_t)h->W * h->N;
    if (ensure_stage(h, n * 6)) return PMS_ERR_CUDA;
    unsigned char *st = reinterpret_cast<unsigned char *>(h->stage);
    CK(cudaMemcpyAsync(st, poses, n * 3 * esz, cudaMemcpyHostToDevice, h->stream));
    if (vels) CK(cudaMemcpyAsync(st + n * 3 * esz, vels, n * 3 * esz, cudaMemcpyHostToDevice, h->stream));
    int rc = state_from_device(h, st, vels ? st + n * 3 * esz : nullptr, esz == 4);
    if (rc) return rc;
    IO_DONE(h);
    return PMS_OK;
}
static int get_ped_state_host(pms_sim *h, void *poses, void *vels, size_t esz)
{
    const size_t n = (size_t)h->W * h->N;
    if (ensure_stage(h, n * 6)) return PMS_ERR_CUDA;
    unsigned char *st = reinterpret_cast<unsigned char *>(h->stage);
    int rc = state_to_device(h, poses ? st : nullptr, vels ? st + n * 3 * esz : nullptr, esz == 4);
    if (rc) return rc;
    if (poses) CK(cudaMemcpyAsync(poses, st, n * 3 * esz, cudaMemcpyDeviceToHost, h->stream));
    if (vels) CK(cudaMemcpyAsync(vels, st + n * 3 * esz, n * 3 * esz, cudaMemcpyDeviceToHost, h->stream));
    IO_DONE(h);
    return PMS_OK;
}

int pms_set_ped_state(pms_sim *h, const double *poses, const double *vels)
{
    HCHECK(h);
    return set_ped_state_host(h, poses, vels, 8);
}
int pms_get_ped_state(pms_sim *h, double *poses, double *vels)
{
    HCHECK(h);
    return get_ped_state_host(h, poses, vels, 8);
}
int pms_set_ped_state_f32(pms_sim *h, const float *poses, const float *vels)
{
    HCHECK(h);
    return set_ped_state_host(h, poses, vels, 4);
}
int pms_get_ped_state_f32(pms_sim *h, float *poses, float *vels)
{
    HCHECK(h);
    return get_ped_state_host(h, poses, vels, 4);
}
int pms_set_ped_state_device(pms_sim *h, const void *poses, const void *vels, int is_f32)
{
    HCHECK(h);
    if (!poses) return fail(PMS_ERR_INVALID, "poses is NULL");
    return state_from_device(h, poses, vels, is_f32 != 0) ? PMS_ERR_CUDA : PMS_OK;
}
int pms_get_ped_state_device(pms_sim *h, void *poses, void *vels, int is_f32)
{
    HCHECK(h);
    return state_to_device(h, poses, vels, is_f32 != 0) ? PMS_ERR_CUDA : PMS_OK;
}

static int upload_converted(pms_sim *h, const double *src, size_t n, void *dst)
{
    if (ensure_stage(h, n)) return PMS_ERR_CUDA;
    CK(cudaMemcpyAsync(h->stage, src, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (h->precision == PMS_FP32) convert_array_kernel<float><<<nblk(n), 256, 0, h->stream>>>(n, h->stage, (float *)dst);
    else convert_array_kernel<double><<<nblk(n), 256, 0, h->stream>>>(n, h->stage, (double *)dst);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return PMS_OK;
}

int pms_set_ped_speeds(pms_sim *h, const double *vmag)
{
    HCHECK(h);
    if (!vmag) return fail(PMS_ERR_INVALID, "vmag is NULL");
    return upload_converted(h, vmag, (size_t)h->W * h->N, h->a.vmag);
}

int pms_set_waypoints_fixed(pms_sim *h, const double *table, int n_rows, const int32_t *start_index, double reach, int loop)
{
    HCHECK(h);
    if (!table || n_rows < 1) return fail(PMS_ERR_INVALID, "table is NULL or n_rows < 1");
    const size_t n = (size_t)h->W * h->N;
    HsfmArgs &a = h->a;
    // validate everything before the handle is touched: a failed call leaves the old table, row count and indices in place
    if (start_index)
        for (size_t k = 0; k < n; ++k)
            if (start_index[k] < 0 || start_index[k] >= n_rows) return fail(PMS_ERR_INVALID, "start_index out of range");
    unsigned char *raw = nullptr;
    CK(dalloc(&raw, n * n_rows * 2 * h->esz));
    int rc = upload_converted(h, table, n * n_rows * 2, raw);
    if (rc) { cudaFree(raw); return rc; }
    if (a.table) cudaFree((void *)a.table);
    a.table = raw;
    h->has_snapshot = false;        // a saved snapshot holds indices into the old table
    if (start_index) {
        CK(cudaMemcpyAsync(a.wp_index, start_index, n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    } else {
        fill_int_kernel<<<nblk(n), 256, 0, h->stream>>>(n, a.wp_index, n_rows > 1 ? 1 : 0);
    }
    a.tracker_kind = TRACKER_FIXED; a.n_table = n_rows; a.reach = reach; a.loop = loop ? 1 : 0;
    if (h->precision == PMS_FP32)
        gather_waypoints_kernel<float><<<nblk(n), 256, 0, h->stream>>>(n, n_rows, (const float *)a.table, a.wp_index, (float *)a.wp);
    else
        gather_waypoints_kernel<double><<<nblk(n), 256, 0, h->stream>>>(n, n_rows, (const double *)a.table, a.wp_index, (double *)a.wp);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    h->waypoints_set = true;
    return PMS_OK;
}

int pms_set_waypoint_queue(pms_sim *h, const double *current, const double *queue, int q_len, double reach)
{
    HCHECK(h);
    if (!current || !queue || q_len < 1) return fail(PMS_ERR_INVALID, "current/queue NULL or q_len < 1");
    const size_t n = (size_t)h->W * h->N;
    HsfmArgs &a = h->a;
    unsigned char *raw = nullptr;
    CK(dalloc(&raw, n * q_len * 2 * h->esz));
    int rc = upload_converted(h, queue, n * q_len * 2, raw);
    if (rc) { cudaFree(raw); return rc; }
    if (a.queue) cudaFree((void *)a.queue);
    a.queue = raw;
    h->has_snapshot = false;        // a saved snapshot points at the old queue allocation and capacity
    rc = upload_converted(h, current, n * 2, a.wp);
    if (rc) return rc;
    CK(cudaMemsetAsync(a.q_head, 0, n * sizeof(int), h->stream));
    fill_int_kernel<<<nblk(n), 256, 0, h->stream>>>(n, a.q_count, q_len);
    CK(cudaMemsetAsync(a.ev_count, 0, sizeof(int), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    a.tracker_kind = TRACKER_QUEUE; a.q_cap = q_len; a.reach = reach;
    h->waypoints_set = true;
    return PMS_OK;
}

int pms_append_waypoints(pms_sim *h, int n, const int32_t *world_ped, const double *points)
{
    HCHECK(h);
    if (h->a.tracker_kind != TRACKER_QUEUE) return fail(PMS_ERR_STATE, "waypoint queue is not configured");
    if (n <= 0) return PMS_OK;
    if (!world_ped || !points) return fail(PMS_ERR_INVALID, "world_ped/points NULL");
    for (int k = 0; k < n; ++k)
        if (world_ped[2 * k] < 0 || world_ped[2 * k] >= h->W || world_ped[2 * k + 1] < 0 || world_ped[2 * k + 1] >= h->N)
            return fail(PMS_ERR_INVALID, "world/ped index out of range");
    if (ensure_stage(h, (size_t)n * 3 + 2)) return PMS_ERR_CUDA;
    int *wp_dev = reinterpret_cast<int *>(h->stage + (size_t)n * 2);
    CK(cudaMemcpyAsync(h->stage, points, (size_t)n * 2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(wp_dev, world_ped, (size_t)n * 2 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->overflow_dev, 0, sizeof(int), h->stream));
    if (h->precision == PMS_FP32)
        append_queue_kernel<float><<<1, 1, 0, h->stream>>>(n, wp_dev, h->stage, h->N, h->a.q_cap, (float *)h->a.queue, h->a.q_head, h->a.q_count, h->overflow_dev);
    else
        append_queue_kernel<double><<<1, 1, 0, h->stream>>>(n, wp_dev, h->stage, h->N, h->a.q_cap, (double *)h->a.queue, h->a.q_head, h->a.q_count, h->overflow_dev);
    int ov = 0;
    CK(cudaMemcpyAsync(&ov, h->overflow_dev, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (ov) return fail(PMS_ERR_STATE, "waypoint queue overflow");
    return PMS_OK;
}

int pms_pop_waypoint_events(pms_sim *h, int32_t *events, int cap, int *n_out)
{
    HCHECK(h);
    int cnt = 0;
    CK(cudaMemcpyAsync(&cnt, h->a.ev_count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (cnt > h->a.ev_cap) return fail(PMS_ERR_STATE, "waypoint event log overflowed (too many steps per call)");
    const int n = cnt < cap ? cnt : cap;
    if (n > 0 && events) CK(cudaMemcpyAsync(events, h->a.events, (size_t)n * 3 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemsetAsync(h->a.ev_count, 0, sizeof(int), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (n_out) *n_out = n;
    if (cnt > cap) return fail(PMS_ERR_INVALID, "event buffer too small");
    return PMS_OK;
}

int pms_get_waypoints(pms_sim *h, double *current, int32_t *index)
{
    HCHECK(h);
    const size_t n = (size_t)h->W * h->N;
    if (current) {
        if (ensure_stage(h, n * 2)) return PMS_ERR_CUDA;
        if (h->precision == PMS_FP32) convert_back_kernel<float><<<nblk(n * 2), 256, 0, h->stream>>>(n * 2, (const float *)h->a.wp, h->stage);
        else convert_back_kernel<double><<<nblk(n * 2), 256, 0, h->stream>>>(n * 2, (const double *)h->a.wp, h->stage);
        CK(cudaMemcpyAsync(current, h->stage, n * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    if (index) CK(cudaMemcpyAsync(index, h->a.wp_index, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return PMS_OK;
}

int pms_set_robot(pms_sim *h, int model, const double *pose, const double *velocity, const double *control,
                  double radius, double wheel_base, int visible)
{
    HCHECK(h);
    if (model < PMS_ROBOT_NONE || model > PMS_ROBOT_BICYCLE) return fail(PMS_ERR_INVALID, "unknown robot model");
    HsfmArgs &a = h->a;
    a.robot_model = model;
    a.robot_visible = visible ? 1 : 0;
    if (model == PMS_ROBOT_NONE) return PMS_OK;
    if (!pose) return fail(PMS_ERR_INVALID, "pose is NULL");
    if (!(radius > 0.0)) return fail(PMS_ERR_INVALID, "robot radius must be > 0");   // core/_motion.py:15
    if (model == PMS_ROBOT_BICYCLE && !(wheel_base > 0.0)) return fail(PMS_ERR_INVALID, "wheel_base must be > 0");
    a.robot_radius = radius; a.wheel_base = wheel_base;
    const int W = h->W;
    std::vector<double> vel((size_t)W * 3, 0.0), ctrl((size_t)W * 2, 0.0), rear((size_t)W * 3, 0.0);
    if (control) std::memcpy(ctrl.data(), control, sizeof(double) * W * 2);
    for (int w = 0; w < W; ++w) {
        const double th = pose[3 * w + 2];
        if (model == PMS_ROBOT_EXTERNAL && velocity) {
            vel[3 * w] = velocity[3 * w]; vel[3 * w + 1] = velocity[3 * w + 1]; vel[3 * w + 2] = velocity[3 * w + 2];
        } else if (model == PMS_ROBOT_UNICYCLE) {   // UnicycleRobotModelState.velocity
            vel[3 * w] = ctrl[2 * w] * std::cos(th); vel[3 * w + 1] = ctrl[2 * w] * std::sin(th); vel[3 * w + 2] = ctrl[2 * w + 1];
        }
        // BicycleRobotModel.__init__: rear-axle pose (robot/_bicycle.py:58-62)
        rear[3 * w] = pose[3 * w] - (wheel_base / 2.0) * std::cos(th);
        rear[3 * w + 1] = pose[3 * w + 1] - (wheel_base / 2.0) * std::sin(th);
        rear[3 * w + 2] = th;
    }
    CK(cudaMemcpyAsync(a.robot_pose, pose, sizeof(double) * W * 3, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(a.robot_vel, vel.data(), sizeof(double) * W * 3, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(a.robot_ctrl, ctrl.data(), sizeof(double) * W * 2, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(a.robot_rear, rear.data(), sizeof(double) * W * 3, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(a.robot_world_collision, 0, sizeof(int) * W, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return PMS_OK;
}

int pms_set_robot_controls(pms_sim *h, const double *controls, int n_rows)
{
    HCHECK(h);
    if (!controls || n_rows <= 0) { h->a.controls = nullptr; h->a.n_controls = 0; return PMS_OK; }
    const size_t n = (size_t)n_rows * h->W * 2;
    if (n > h->controls_cap) {
        if (h->controls_dev) cudaFree(h->controls_dev);
        h->controls_dev = nullptr; h->controls_cap = 0;
        CK(cudaMalloc((void **)&h->controls_dev, n * sizeof(double)));
        h->controls_cap = n;
    }
    CK(cudaMemcpyAsync(h->controls_dev, controls, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->a.controls = h->controls_dev; h->a.n_controls = n_rows;
    return PMS_OK;
}

int pms_set_robot_controls_device(pms_sim *h, const double *controls_dev, int n_rows)
{
    HCHECK(h);
    if (!controls_dev || n_rows <= 0) { h->a.controls = nullptr; h->a.n_controls = 0; return PMS_OK; }
    h->a.controls = controls_dev; h->a.n_controls = n_rows;      // read in place by the launches that follow: no copy, no synchronisation
    return PMS_OK;
}

int pms_get_robot(pms_sim *h, double *pose, double *velocity, int32_t *world_collision)
{
    HCHECK(h);
    const int W = h->W;
    if (pose) CK(cudaMemcpyAsync(pose, h->a.robot_pose, sizeof(double) * W * 3, cudaMemcpyDeviceToHost, h->stream));
    if (velocity) CK(cudaMemcpyAsync(velocity, h->a.robot_vel, sizeof(double) * W * 3, cudaMemcpyDeviceToHost, h->stream));
    if (world_collision) CK(cudaMemcpyAsync(world_collision, h->a.robot_world_collision, sizeof(int) * W, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return PMS_OK;
}

int pms_set_map(pms_sim *h, int kind, const double *data, int count, int per_world)
{
    HCHECK(h);
    if (kind < PMS_MAP_EMPTY || kind > PMS_MAP_AABB) return fail(PMS_ERR_INVALID, "unknown map kind");
    HsfmArgs &a = h->a;
    if (kind == PMS_MAP_EMPTY || count <= 0) { a.map_kind = kind; a.map_count = 0; a.map_data = nullptr; a.map_stride = 0; return PMS_OK; }
    if (!data) return fail(PMS_ERR_INVALID, "map data is NULL");
    const int k = kind == PMS_MAP_CIRCLES ? 3 : 4;
    const size_t n = (size_t)count * k * (per_world ? h->W : 1);
    if (n > h->map_cap) {
        if (h->map_dev) cudaFree(h->map_dev);
        h->map_dev = nullptr; h->map_cap = 0;
        CK(cudaMalloc((void **)&h->map_dev, n * sizeof(double)));
        h->map_cap = n;
    }
    CK(cudaMemcpyAsync(h->map_dev, data, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    a.map_kind = kind; a.map_count = count; a.map_data = h->map_dev; a.map_stride = per_world ? count * k : 0;
    return PMS_OK;
}

int pms_set_accel_noise(pms_sim *h, const double *noise)
{
    HCHECK(h);
    const size_t n = (size_t)h->W * h->N * 2;
    if (!noise) { h->a.accel_noise = nullptr; return PMS_OK; }
    h->a.accel_noise_std = 0.0;            // an explicit noise array replaces the device generator
    if (!h->noise_dev) CK(cudaMalloc((void **)&h->noise_dev, n * sizeof(double)));
    CK(cudaMemcpyAsync(h->noise_dev, noise, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->a.accel_noise = h->noise_dev;
    return PMS_OK;
}

int pms_set_accel_noise_std(pms_sim *h, double noise_std, uint64_t seed)
{
    HCHECK(h);
    if (!(noise_std >= 0.0)) return fail(PMS_ERR_INVALID, "noise_std must be >= 0");
    h->a.accel_noise_std = noise_std; h->a.noise_seed = seed;
    if (noise_std > 0.0) h->a.accel_noise = nullptr;
    return PMS_OK;
}

int pms_set_lidar_noise(pms_sim *h, int enabled, double drop_prob, const double *point_mean, const double *point_cov, uint64_t seed)
{
    HCHECK(h);
    if (!enabled) { h->lidar_noise.enabled = 0; return PMS_OK; }
    if (!(drop_prob >= 0.0 && drop_prob <= 1.0)) return fail(PMS_ERR_INVALID, "need 0 <= drop_prob <= 1");          // _lidar.py:62
    const double mx = point_mean ? point_mean[0] : 0.0, my = point_mean ? point_mean[1] : 0.0;
    const double c00 = point_cov ? point_cov[0] : 0.001, c01 = point_cov ? point_cov[1] : 0.0,
                 c10 = point_cov ? point_cov[2] : 0.0, c11 = point_cov ? point_cov[3] : 0.001;              // :57
    if (!(c00 >= 0.0 && c01 >= 0.0 && c10 >= 0.0 && c11 >= 0.0)) return fail(PMS_ERR_INVALID, "point_cov must be >= 0"); // :61
    if (c01 != c10) return fail(PMS_ERR_INVALID, "point_cov must be symmetric");
    const double l00 = std::sqrt(c00), l10 = l00 > 0.0 ? c10 / l00 : 0.0, rest = c11 - l10 * l10;
    if (rest < -1e-12 * (c11 > 1.0 ? c11 : 1.0)) return fail(PMS_ERR_INVALID, "point_cov must be positive semi-definite");
    h->lidar_noise = LidarNoise{drop_prob, mx, my, l00, l10, rest > 0.0 ? std::sqrt(rest) : 0.0, seed, 1};
    return PMS_OK;
}

int pms_set_omni_noise(pms_sim *h, int enabled, double misdetection_prob, double distance_mu, double distance_sigma, uint64_t seed)
{
    HCHECK(h);
    if (!enabled) { h->omni_noise.enabled = 0; return PMS_OK; }
    if (!(misdetection_prob <= 1.0)) return fail(PMS_ERR_INVALID, "misdetection_prob must be <= 1");   // _omni_obstacle_detector.py:49
    if (!(distance_sigma >= 0.0)) return fail(PMS_ERR_INVALID, "sigma must be >= 0");
    h->omni_noise = OmniNoise{misdetection_prob, distance_mu, distance_sigma, seed, 1};
    return PMS_OK;
}

void pms_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    const uint4 r = philox4x32_10(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), make_uint2(key[0], key[1]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// The sensor pass of one tick for a pedestrians model that runs elsewhere (ORCA handle, replayed trajectories, a custom
// model): poses + robot in, collision list + detector reading out, one call and one host synchronisation.
int pms_sense_tick(pms_sim *h, const double *poses, const double *robot_pose, double robot_radius, uint32_t *coll_mask,
                   uint32_t *det_mask, double *det_vals)
{
    HCHECK(h);
    if (!poses || !robot_pose) return fail(PMS_ERR_INVALID, "poses / robot_pose is NULL");
    if (!(robot_radius > 0.0)) return fail(PMS_ERR_INVALID, "robot radius must be > 0");
    if (h->det_noise.enabled) return fail(PMS_ERR_STATE, "pms_sense_tick returns the exact reading: switch the detector noise off");
    HsfmArgs &a = h->a;
    const int W = h->W;
    const size_t n = (size_t)W * h->N, nm = (size_t)W * h->words;
    // host block: poses (3n doubles) | robot pose (3W) -- then reused for the results: det_vals (2n doubles) | coll | det
    const size_t in_bytes = (n * 3 + (size_t)W * 3) * 8, o_coll = n * 16, o_det = o_coll + nm * 4, out_bytes = o_det + nm * 4;
    const size_t o_out = (in_bytes + 15) & ~(size_t)15;      // double2 records: 16-byte aligned
    const bool direct = o_out + out_bytes <= 262144;         // small tick: one kernel on the mapped host block (outputs behind the inputs)
    const size_t need = direct ? o_out + out_bytes : (in_bytes > out_bytes ? in_bytes : out_bytes);
    if (need > h->tick_bytes) {
        if (h->tick_dev) cudaFree(h->tick_dev);
        if (h->tick_host) cudaFreeHost(h->tick_host);
        h->tick_dev = h->tick_host = nullptr; h->tick_bytes = 0;
        CK(cudaMalloc((void **)&h->tick_dev, need));
        CK(cudaMallocHost((void **)&h->tick_host, need));
        h->tick_bytes = need;
    }
    if (ensure_stage(h, n * 6)) return PMS_ERR_CUDA;
    double *st = reinterpret_cast<double *>(h->tick_host);
    std::memcpy(st, poses, n * 24);
    std::memcpy(st + n * 3, robot_pose, (size_t)W * 24);
    a.robot_model = PMS_ROBOT_EXTERNAL; a.robot_radius = robot_radius;
    if (direct) {
        fill_consts(a);
        const dim3 grid(W, (h->Np + 255) / 256);
        const int threads = h->Np < 256 ? h->Np : 256;
        unsigned char *o = h->tick_host + o_out;
        if (h->precision == PMS_FP32)
            sense_tick_kernel<float><<<grid, threads, 0, h->stream>>>(a, st, st + n * 3, (double2 *)o, (uint32_t *)(o + o_coll), (uint32_t *)(o + o_det), h->compensated);
        else
            sense_tick_kernel<double><<<grid, threads, 0, h->stream>>>(a, st, st + n * 3, (double2 *)o, (uint32_t *)(o + o_coll), (uint32_t *)(o + o_det), false);
        CK(cudaGetLastError());
        h->large.sorted_valid = false;
        h->info.kernel_launches++;
        CK(cudaStreamSynchronize(h->stream));
        if (det_vals) std::memcpy(det_vals, o, n * 16);
        if (coll_mask) std::memcpy(coll_mask, o + o_coll, nm * 4);
        if (det_mask) std::memcpy(det_mask, o + o_det, nm * 4);
        return PMS_OK;
    }
    CK(cudaMemcpyAsync(h->stage, st, n * 24, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->stage + n * 3, 0, n * 24, h->stream));
    CK(cudaMemcpyAsync(a.robot_pose, st + n * 3, (size_t)W * 24, cudaMemcpyHostToDevice, h->stream));
    if (state_from_device(h, h->stage, h->stage + n * 3, false)) return PMS_ERR_CUDA;
    int rc = clear_sensor_outputs(h);
    if (rc) return rc;
    fill_consts(a);
    const dim3 grid(W, (h->Np + 255) / 256);
    const int threads = h->Np < 256 ? h->Np : 256;
    if (h->precision == PMS_FP32) sense_kernel<float><<<grid, threads, 0, h->stream>>>(a);
    else sense_kernel<double><<<grid, threads, 0, h->stream>>>(a);
    h->info.kernel_launches++;
    unsigned char *d = h->tick_dev;
    CK(cudaMemcpyAsync(d, a.det_vals, n * 16, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(d + o_coll, a.coll_mask, nm * 4, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(d + o_det, a.det_mask, nm * 4, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->tick_host, d, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const unsigned char *q = h->tick_host;
    if (det_vals) std::memcpy(det_vals, q, n * 16);
    if (coll_mask) std::memcpy(coll_mask, q + o_coll, nm * 4);
    if (det_mask) std::memcpy(det_mask, q + o_det, nm * 4);
    return PMS_OK;
}

int pms_detector_config(pms_sim *h, int enabled, double max_dist, double fov, int return_type)
{
    HCHECK(h);
    if (return_type < PMS_DET_POLAR || return_type > PMS_DET_ABSOLUTE)
        return fail(PMS_ERR_INVALID, "unknown detector return type");   // _pedestrian_detector.py:29-32
    h->a.det_enabled = enabled ? 1 : 0; h->a.det_max_dist = max_dist; h->a.det_half_fov = fov / 2.0;
    h->det_noise.return_type = return_type;
    h->a.det_type = h->det_noise.enabled ? (int)DET_POLAR : return_type;    // noise is added in the polar frame
    return PMS_OK;
}

int pms_set_detector_noise(pms_sim *h, int enabled, double misdetection_prob, double distance_mu, double distance_sigma,
                           double angle_mu, double angle_sigma, uint64_t seed)
{
    HCHECK(h);
    if (enabled) {
        if (!(misdetection_prob <= 1.0)) return fail(PMS_ERR_INVALID, "misdetection_prob must be <= 1");   // _pedestrian_detector.py:69
        if (!(distance_sigma >= 0.0) || !(angle_sigma >= 0.0)) return fail(PMS_ERR_INVALID, "sigma must be >= 0");
        if (!h->noisy_mask) {
            CK(cudaMalloc((void **)&h->noisy_mask, sizeof(uint32_t) * (size_t)h->W * h->words));
            CK(cudaMalloc((void **)&h->noisy_vals, sizeof(double2) * (size_t)h->W * h->N));
        }
    }
    const int user_type = h->det_noise.enabled ? h->det_noise.return_type : h->a.det_type;
    h->det_noise = DetNoise{misdetection_prob, distance_mu, distance_sigma, angle_mu, angle_sigma, seed, enabled ? 1 : 0, user_type};
    h->a.det_type = enabled ? (int)DET_POLAR : user_type;
    return PMS_OK;
}

int pms_hsfm_step_async(pms_sim *h, double dt, int n_steps)
{
    HCHECK(h);
    if (n_steps < 0) return fail(PMS_ERR_INVALID, "n_steps < 0");
    if (n_steps == 0) return PMS_OK;
    if (!h->waypoints_set) return fail(PMS_ERR_STATE, "Waypoint tracker is not initialized");
    h->a.dt = dt; h->a.n_steps = n_steps;
    fill_consts(h->a);
    int rc;
    // j_split < 0 selects the directed-pair kernel for small worlds too (tests, comparisons)
    const bool warp_ok = h->precision != PMS_FP64 && h->Np <= 32 * WARP_MAX_TILES && h->tune_S >= 0;
    // FP32 worlds of 512 pedestrians and more take the pruned large-crowd path (sorted copy + chunk kernel) although they would
    // fit the directed-pair kernel: measured 64 x 992: 77 -> 18 us per step, 256 x 512: 50 -> 30 (bench crowd, 100-step calls);
    // below 512 the all-pairs kernel wins when the world is dense (nothing to prune).  j_split < 0 keeps the directed-pair kernel.
    const bool mid_fp32 = h->precision == PMS_FP32 && h->N >= 512 && !h->no_prune && !h->no_reorder && h->tune_S >= 0 && h->tune_cluster == 0;
    const bool large_path = h->force_large || (!warp_ok && (h->Np + 32 > 1024 || mid_fp32));
    // the large-crowd kernel sets mask bits with atomicOr when it works on the sorted copy: start from zero
    if ((h->a.robot_model == ROBOT_NONE || large_path) && (rc = clear_sensor_outputs(h))) return rc;
    if ((rc = prepare_sensor_plan(h, n_steps, large_path))) return rc;
    if (large_path) rc = launch_large(h);
    else if (warp_ok && duo_eligible(h) && duo_preferred(h)) rc = launch_duo(h);
    else if (warp_ok) rc = launch_warp(h);
    else rc = launch_small(h);
    if (rc) return rc;
    if (h->precision == PMS_FP32 && !h->compensated && (large_path || !warp_ok)) {   // plain fp32 state: low parts are gone
        const size_t n = (size_t)h->W * h->N;
        CK(cudaMemsetAsync(h->a.pos_lo, 0, n * 16, h->stream));
        CK(cudaMemsetAsync(h->a.th_lo, 0, n * 8, h->stream));
    }
    h->a.steps_done += n_steps;
    return PMS_OK;
}

int pms_sync(pms_sim *h)
{
    if (!h) return fail(PMS_ERR_INVALID, "handle is NULL");
    Guard gd(h->device);
    CK(cudaStreamSynchronize(h->stream));
    return PMS_OK;
}

int pms_hsfm_step(pms_sim *h, double dt, int n_steps)
{
    int rc = pms_hsfm_step_async(h, dt, n_steps);
    if (rc) return rc;
    return pms_sync(h);
}

void *pms_stream(pms_sim *h) { return h ? (void *)h->stream : nullptr; }

// One tick of the drop-in classes in one call and ONE host synchronisation: robot upload (pinned staging, asynchronous),
// n_steps fused steps, every output gathered into one device block, one device-to-host copy.
int pms_step_tick(pms_sim *h, double dt, int n_steps, const double *robot_pose, const double *robot_velocity, double robot_radius,
                  int visible, double *poses, double *vels, int32_t *wp_index, uint32_t *coll_mask, uint32_t *det_mask,
                  double *det_vals, int32_t *first_nonfinite_step, int32_t *n_events)
{
    HCHECK(h);
    if (n_steps < 1) return fail(PMS_ERR_INVALID, "n_steps < 1");
    HsfmArgs &a = h->a;
    const int W = h->W;
    const size_t n = (size_t)W * h->N, nm = (size_t)W * h->words;
    // layout of the tick block: poses | vels | det_vals (doubles) | wp_index | nonfinite | n_events (ints) | coll | det (u32);
    // the head of the host block doubles as the staging area of the robot upload (pose | velocity | control | rear: 11 W doubles, W ints)
    const size_t o_pose = 0, o_vel = o_pose + n * 24, o_vals = o_vel + n * 24, o_idx = o_vals + n * 16, o_nf = o_idx + n * 4,
                 o_ev = o_nf + (size_t)W * 4, o_coll = o_ev + 4, o_det = o_coll + nm * 4, total = o_det + nm * 4;
    const size_t need = total > (size_t)W * 96 ? total : (size_t)W * 96;
    if (need > h->tick_bytes) {
        if (h->tick_dev) cudaFree(h->tick_dev);
        if (h->tick_host) cudaFreeHost(h->tick_host);
        h->tick_dev = h->tick_host = nullptr; h->tick_bytes = 0;
        CK(cudaMalloc((void **)&h->tick_dev, need));
        CK(cudaMallocHost((void **)&h->tick_host, need));
        h->tick_bytes = need;
    }
    if (!robot_pose) {
        a.robot_model = PMS_ROBOT_NONE;
    } else {
        if (!(robot_radius > 0.0)) return fail(PMS_ERR_INVALID, "robot radius must be > 0");
        a.robot_model = PMS_ROBOT_EXTERNAL; a.robot_visible = visible ? 1 : 0; a.robot_radius = robot_radius;
        double *st = reinterpret_cast<double *>(h->tick_host);       // the previous tick was synchronised: the block is free
        double *s_pose = st, *s_vel = st + 3 * W, *s_ctrl = st + 6 * W, *s_rear = st + 8 * W;
        for (int w = 0; w < W; ++w) {
            const double th = robot_pose[3 * w + 2];
            for (int k = 0; k < 3; ++k) { s_pose[3 * w + k] = robot_pose[3 * w + k]; s_vel[3 * w + k] = robot_velocity ? robot_velocity[3 * w + k] : 0.0; }
            s_ctrl[2 * w] = s_ctrl[2 * w + 1] = 0.0;
            s_rear[3 * w] = robot_pose[3 * w] - (a.wheel_base / 2.0) * std::cos(th);
            s_rear[3 * w + 1] = robot_pose[3 * w + 1] - (a.wheel_base / 2.0) * std::sin(th);
            s_rear[3 * w + 2] = th;
        }
        std::memset(st + 11 * W, 0, sizeof(int) * W);                 // world-collision flags
        // the device block has the same layout: one copy instead of four copies and a memset
        CK(cudaMemcpyAsync(h->robot_block, st, sizeof(double) * W * 11 + sizeof(int) * W, cudaMemcpyHostToDevice, h->stream));
    }
    int rc = pms_hsfm_step_async(h, dt, n_steps);
    if (rc) return rc;
    // small ticks (the drop-in classes: one world): the gather kernel writes straight into the pinned host block (mapped, UVA)
    // -- no device-to-host copy in the chain; large ones go through the device block and one copy
    const bool direct = total <= 65536;
    unsigned char *d = direct ? h->tick_host : h->tick_dev;
    {
        const unsigned nb_ = nblk(n > nm ? n : nm);
        if (h->precision == PMS_FP32)
            tick_pack_kernel<float><<<nb_, 256, 0, h->stream>>>(n, W, h->words, a, (double *)(d + o_pose), (double *)(d + o_vel), (double2 *)(d + o_vals),
                                                                (int *)(d + o_idx), (uint32_t *)(d + o_coll), (uint32_t *)(d + o_det), (int *)(d + o_nf), (int *)(d + o_ev));
        else
            tick_pack_kernel<double><<<nb_, 256, 0, h->stream>>>(n, W, h->words, a, (double *)(d + o_pose), (double *)(d + o_vel), (double2 *)(d + o_vals),
                                                                 (int *)(d + o_idx), (uint32_t *)(d + o_coll), (uint32_t *)(d + o_det), (int *)(d + o_nf), (int *)(d + o_ev));
    }
    CK(cudaGetLastError());
    if (!direct) CK(cudaMemcpyAsync(h->tick_host, d, total, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const unsigned char *q = h->tick_host;
    if (poses) std::memcpy(poses, q + o_pose, n * 24);
    if (vels) std::memcpy(vels, q + o_vel, n * 24);
    if (det_vals) std::memcpy(det_vals, q + o_vals, n * 16);
    if (wp_index) std::memcpy(wp_index, q + o_idx, n * 4);
    if (coll_mask) std::memcpy(coll_mask, q + o_coll, nm * 4);
    if (det_mask) std::memcpy(det_mask, q + o_det, nm * 4);
    if (n_events) std::memcpy(n_events, q + o_ev, 4);
    if (first_nonfinite_step) {
        std::memcpy(first_nonfinite_step, q + o_nf, (size_t)W * 4);
        for (int w = 0; w < W; ++w) if (first_nonfinite_step[w] == 0x7fffffff) first_nonfinite_step[w] = 0;
    }
    return PMS_OK;
}

int pms_sense(pms_sim *h)
{
    HCHECK(h);
    int rc = clear_sensor_outputs(h);
    if (rc) return rc;
    fill_consts(h->a);
    if (h->a.robot_model != ROBOT_NONE) {
        const dim3 grid(h->W, (h->Np + 255) / 256);
        const int threads = h->Np < 256 ? h->Np : 256;
        if (h->precision == PMS_FP32) sense_kernel<float><<<grid, threads, 0, h->stream>>>(h->a);
        else sense_kernel<double><<<grid, threads, 0, h->stream>>>(h->a);
        CK(cudaGetLastError());
        h->info.kernel_launches++;
    }
    CK(cudaStreamSynchronize(h->stream));
    return PMS_OK;
}

int pms_get_detector(pms_sim *h, uint32_t *mask, double *values)
{
    HCHECK(h);
    const size_t n = (size_t)h->W * h->N;
    const uint32_t *src_mask = h->a.det_mask;
    const double2 *src_vals = h->a.det_vals;
    if (h->det_noise.enabled && h->a.det_enabled) {     // the stored reading stays exact; the noisy view is made per call
        detector_noise_kernel<<<dim3(h->W, (h->Np + 255) / 256), h->Np < 256 ? h->Np : 256, 0, h->stream>>>(h->W, h->N, h->words, h->a.det_mask, h->a.det_vals, h->noisy_mask,
                                                             h->noisy_vals, h->a.robot_pose, h->det_noise, h->a.steps_done);
        CK(cudaGetLastError());
        src_mask = h->noisy_mask; src_vals = h->noisy_vals;
    }
    if (mask) CK(cudaMemcpyAsync(mask, src_mask, sizeof(uint32_t) * h->W * h->words, cudaMemcpyDeviceToHost, h->stream));
    if (values) CK(cudaMemcpyAsync(values, src_vals, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, h->stream));
    IO_DONE(h);
    return PMS_OK;
}

int pms_get_collisions(pms_sim *h, uint32_t *mask)
{
    HCHECK(h);
    if (mask) CK(cudaMemcpyAsync(mask, h->a.coll_mask, sizeof(uint32_t) * h->W * h->words, cudaMemcpyDeviceToHost, h->stream));
    IO_DONE(h);
    return PMS_OK;
}

int pms_get_detector_f32(pms_sim *h, uint32_t *mask, float *values)
{
    HCHECK(h);
    if (h->det_noise.enabled && h->a.det_enabled) return fail(PMS_ERR_UNSUPPORTED, "pms_get_detector_f32 returns the exact reading: use pms_get_detector with the device noise");
    const size_t n = (size_t)h->W * h->N;
    if (mask) CK(cudaMemcpyAsync(mask, h->a.det_mask, sizeof(uint32_t) * h->W * h->words, cudaMemcpyDeviceToHost, h->stream));
    if (values) {
        if (ensure_stage(h, n)) return PMS_ERR_CUDA;
        convert_vals_kernel<float><<<nblk(n), 256, 0, h->stream>>>(n, h->a.det_vals, reinterpret_cast<float *>(h->stage));
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(values, h->stage, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, h->stream));
    }
    IO_DONE(h);
    return PMS_OK;
}

int pms_device_buffer(pms_sim *h, int which, void **ptr, uint64_t *bytes)
{
    HCHECK(h);
    if (!ptr) return fail(PMS_ERR_INVALID, "ptr is NULL");
    const HsfmArgs &a = h->a;
    const pms_sim::Sched &sc = h->sched;
    const uint64_t W = h->W, N = h->N, words = h->words, nm = W * words, n = W * N;
    void *p = nullptr;
    uint64_t b = 0;
    switch (which) {
    case PMS_BUF_DET_MASK: p = a.det_mask; b = nm * 4; break;
    case PMS_BUF_COLL_MASK: p = a.coll_mask; b = nm * 4; break;
    case PMS_BUF_DET_VALUES: p = a.det_vals; b = n * 16; break;
    case PMS_BUF_ROBOT_POSE: p = a.robot_pose; b = W * 24; break;
    case PMS_BUF_ROBOT_VELOCITY: p = a.robot_vel; b = W * 24; break;
    case PMS_BUF_DETECTIONS: p = a.det_count; b = W * 8; break;
    case PMS_BUF_COLLISIONS: p = a.coll_count; b = W * 8; break;
    case PMS_BUF_FIRST_NONFINITE: p = a.nonfinite; b = W * 4; break;
    case PMS_BUF_WAYPOINT_INDEX: p = a.wp_index; b = n * 4; break;
    case PMS_BUF_REC_DET_MASK: p = sc.det_mask; b = (uint64_t)sc.n_fires[SENSOR_DETECTOR] * nm * 4; break;
    case PMS_BUF_REC_DET_VALUES: p = (sc.flags[SENSOR_DETECTOR] & PMS_SENSOR_VALUES) ? sc.det_vals : nullptr; b = (uint64_t)sc.n_fires[SENSOR_DETECTOR] * n * 16; break;
    case PMS_BUF_REC_COLL_MASK: p = sc.coll_mask; b = (uint64_t)sc.n_fires[SENSOR_COLLISIONS] * nm * 4; break;
    case PMS_BUF_REC_LIDAR_HIT: p = sc.lidar_hit; b = (uint64_t)sc.n_fires[SENSOR_LIDAR] * W * sc.lidar_beams * 4; break;
    case PMS_BUF_REC_LIDAR_POINTS: p = sc.lidar_pts; b = (uint64_t)sc.n_fires[SENSOR_LIDAR] * W * sc.lidar_beams * 16; break;
    case PMS_BUF_REC_OMNI: p = sc.omni; b = (uint64_t)sc.n_fires[SENSOR_OMNI] * W * 8; break;
    default: return fail(PMS_ERR_INVALID, "unknown buffer");
    }
    if (!p) b = 0;
    *ptr = p;
    if (bytes) *bytes = b;
    return PMS_OK;
}

int pms_get_stats(pms_sim *h, int64_t *detections, int64_t *collisions, int32_t *first_nonfinite_step)
{
    HCHECK(h);
    const int W = h->W;
    if (detections) CK(cudaMemcpyAsync(detections, h->a.det_count, sizeof(int64_t) * W, cudaMemcpyDeviceToHost, h->stream));
    if (collisions) CK(cudaMemcpyAsync(collisions, h->a.coll_count, sizeof(int64_t) * W, cudaMemcpyDeviceToHost, h->stream));
    if (first_nonfinite_step) CK(cudaMemcpyAsync(first_nonfinite_step, h->a.nonfinite, sizeof(int) * W, cudaMemcpyDeviceToHost, h->stream));
    if (h->async_io) return PMS_OK;       // the caller maps the "never" sentinel INT32_MAX to 0 after pms_sync
    CK(cudaStreamSynchronize(h->stream));
    if (first_nonfinite_step)
        for (int w = 0; w < W; ++w) if (first_nonfinite_step[w] == 0x7fffffff) first_nonfinite_step[w] = 0;
    return PMS_OK;
}

int pms_reset_stats(pms_sim *h)
{
    HCHECK(h);
    CK(cudaMemsetAsync(h->a.det_count, 0, sizeof(int64_t) * h->W, h->stream));
    CK(cudaMemsetAsync(h->a.coll_count, 0, sizeof(int64_t) * h->W, h->stream));
    fill_int_kernel<<<nblk(h->W), 256, 0, h->stream>>>((size_t)h->W, h->a.nonfinite, 0x7fffffff);
    CK(cudaStreamSynchronize(h->stream));
    h->a.steps_done = 0;
    return PMS_OK;
}

// ---- snapshot ----
static void snapshot_regions(pms_sim *h, std::vector<std::pair<void *, size_t>> &r)
{
    const HsfmArgs &a = h->a;
    const size_t n = (size_t)h->W * h->N, W = h->W;
    r = {{a.pos, n * 16}, {a.vel, n * 16}, {a.th, n * h->esz * 2}, {a.wp, n * h->esz * 2},
         {a.wp_index, n * 4}, {a.q_head, n * 4}, {a.q_count, n * 4},
         {a.robot_pose, W * 24}, {a.robot_vel, W * 24}, {a.robot_ctrl, W * 16}, {a.robot_rear, W * 24},
         {a.robot_world_collision, W * 4}, {a.det_count, W * 8}, {a.coll_count, W * 8}, {a.nonfinite, W * 4}};
    if (a.queue) r.push_back({(void *)a.queue, n * a.q_cap * 2 * h->esz});
    if (a.th_lo) r.push_back({a.th_lo, n * 8});      // (the {x,y,vx,vy} low parts live in a.vel)
}

int pms_snapshot_save(pms_sim *h)
{
    HCHECK(h);
    std::vector<std::pair<void *, size_t>> r;
    snapshot_regions(h, r);
    for (void *p : h->snap_dst) if (p) cudaFree(p);
    h->snap_dst.assign(r.size(), nullptr);
    for (size_t k = 0; k < r.size(); ++k) {
        CK(cudaMalloc(&h->snap_dst[k], r[k].second ? r[k].second : 1));
        CK(cudaMemcpyAsync(h->snap_dst[k], r[k].first, r[k].second, cudaMemcpyDeviceToDevice, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    h->snap_src = r;
    h->snap_steps = h->a.steps_done;
    for (int k = 0; k < SENSOR_KINDS; ++k) h->sched.snap_hold[k] = h->sched.hold_host[k];
    h->snap_replay_frame = h->replay_frame;
    h->has_snapshot = true;
    return PMS_OK;
}

int pms_snapshot_restore(pms_sim *h)
{
    HCHECK(h);
    if (!h->has_snapshot) return fail(PMS_ERR_STATE, "no snapshot saved");
    for (size_t k = 0; k < h->snap_src.size(); ++k)
        CK(cudaMemcpyAsync(h->snap_src[k].first, h->snap_dst[k], h->snap_src[k].second, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemsetAsync(h->a.ev_count, 0, sizeof(int), h->stream));
    h->a.steps_done = h->snap_steps;
    h->large.sorted_valid = false;
    for (int k = 0; k < SENSOR_KINDS; ++k) h->sched.hold_host[k] = h->sched.snap_hold[k];
    h->replay_frame = h->snap_replay_frame;
    if (h->sched.hold_dev)      // snap_hold is a member of the handle: stable until the copy has run
        CK(cudaMemcpyAsync(h->sched.hold_dev, h->sched.snap_hold, sizeof(double) * SENSOR_KINDS, cudaMemcpyHostToDevice, h->stream));
    return PMS_OK;
}

// ---- lidar / omni ----
int pms_lidar_scan(pms_sim *h, int n_beams, double max_dist, double resolution, int32_t *hit_index, double *points)
{
    HCHECK(h);
    if (n_beams <= 0) return fail(PMS_ERR_INVALID, "n_beams must be > 0");                          // _lidar.py:15
    if (!(max_dist > resolution && resolution > 0.0)) return fail(PMS_ERR_INVALID, "need max_dist > resolution > 0"); // :16
    if (h->a.robot_model == ROBOT_NONE) return fail(PMS_ERR_STATE, "no robot configured");
    const size_t nb = (size_t)h->W * n_beams;
    // np.linspace(0, 2*pi, n_beams) and np.arange(0, max_dist, resolution): _lidar.py:111-112
    std::vector<double> dirs;
    lidar_directions(n_beams, dirs);
    const int n_samples = (int)std::ceil((max_dist - 0.0) / resolution);
    // stage layout: dirs | points(double2) | hit(int)
    const size_t need = (size_t)n_beams * 2 + nb * 2 + (nb + 1) / 2 + 2;
    if (ensure_stage(h, need)) return PMS_ERR_CUDA;
    double *d_dirs = h->stage;
    double2 *d_pts = reinterpret_cast<double2 *>(h->stage + (size_t)n_beams * 2);
    int32_t *d_hit = reinterpret_cast<int32_t *>(h->stage + (size_t)n_beams * 2 + nb * 2);
    CK(cudaMemcpyAsync(d_dirs, dirs.data(), dirs.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    LidarArgs la{h->W, n_beams, n_samples, resolution, h->a.robot_pose, d_dirs, h->a.map_kind, h->a.map_count,
                 h->a.map_stride, h->a.map_data, d_hit, d_pts};
    lidar_kernel<<<nblk(nb, 128), 128, 0, h->stream>>>(la);
    CK(cudaGetLastError());
    h->info.kernel_launches++;
    if (h->lidar_noise.enabled) {
        lidar_noise_kernel<<<nblk(nb, 128), 128, 0, h->stream>>>(nb, d_hit, d_pts, h->lidar_noise, h->a.steps_done);
        CK(cudaGetLastError());
    }
    if (hit_index) CK(cudaMemcpyAsync(hit_index, d_hit, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (points) CK(cudaMemcpyAsync(points, d_pts, nb * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return PMS_OK;
}

int pms_omni_scan(pms_sim *h, double *min_dist)
{
    HCHECK(h);
    if (h->a.robot_model == ROBOT_NONE) return fail(PMS_ERR_STATE, "no robot configured");
    if (ensure_stage(h, (size_t)h->W)) return PMS_ERR_CUDA;
    omni_kernel<<<nblk(h->W, 128), 128, 0, h->stream>>>(h->W, h->a.robot_pose, h->a.map_kind, h->a.map_count,
                                                        h->a.map_stride, h->a.map_data, h->stage);
    CK(cudaGetLastError());
    h->info.kernel_launches++;
    if (h->omni_noise.enabled) {
        omni_noise_kernel<<<nblk(h->W, 128), 128, 0, h->stream>>>(h->W, h->stage, h->omni_noise, h->a.steps_done);
        CK(cudaGetLastError());
    }
    if (min_dist) CK(cudaMemcpyAsync(min_dist, h->stage, sizeof(double) * h->W, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return PMS_OK;
}

// ---- recorded trajectories: ReplayPedestriansPolicy as a batched gather ----
int pms_set_replay(pms_sim *h, const double *poses, const double *vels, int n_frames, int loop)
{
    HCHECK(h);
    if (!poses || !vels || n_frames < 1) return fail(PMS_ERR_INVALID, "poses / vels NULL or n_frames < 1");
    const size_t per = (size_t)h->W * h->N * 3, total = per * n_frames;
    CK(cudaStreamSynchronize(h->stream));
    if (h->replay_poses) cudaFree(h->replay_poses);
    if (h->replay_vels) cudaFree(h->replay_vels);
    h->replay_poses = h->replay_vels = nullptr; h->replay_frames = 0;
    CK(cudaMalloc((void **)&h->replay_poses, total * sizeof(double)));
    CK(cudaMalloc((void **)&h->replay_vels, total * sizeof(double)));
    CK(cudaMemcpy(h->replay_poses, poses, total * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->replay_vels, vels, total * sizeof(double), cudaMemcpyHostToDevice));
    h->replay_frames = n_frames; h->replay_loop = loop ? 1 : 0; h->replay_frame = 0;
    if (state_from_device(h, h->replay_poses, h->replay_vels, false)) return PMS_ERR_CUDA;     // frame 0 is the state (_replay.py:36-38)
    CK(cudaStreamSynchronize(h->stream));
    return PMS_OK;
}

static int replay_step_async(pms_sim *h, double dt, int n_steps)
{
    if (!h->replay_frames) return fail(PMS_ERR_STATE, "no recorded trajectories: call pms_set_replay first");
    if (n_steps < 0) return fail(PMS_ERR_INVALID, "n_steps < 0");
    if (n_steps == 0) return PMS_OK;
    h->a.dt = dt; h->a.n_steps = n_steps;
    fill_consts(h->a);
    int rc;
    if (h->a.robot_model == ROBOT_NONE && (rc = clear_sensor_outputs(h))) return rc;
    if ((rc = prepare_sensor_plan(h, n_steps, false))) return rc;
    const ReplayArgs ra{h->replay_poses, h->replay_frames, h->replay_loop, h->replay_frame};
    CK(replay_launch(h->a, ra, h->a.sp.on != 0, h->stream));
    h->info.kernel = 5; h->info.threads_per_block = 128; h->info.blocks = (h->W + 3) / 4; h->info.worlds_per_block = 4;
    h->info.j_split = 0; h->info.smem_bytes = 0; h->info.n_chunks = 1; h->info.blocks_per_sm = 0;
    h->info.kernel_launches++;
    h->replay_frame = replay_advance(h->replay_frame, h->replay_frames, h->replay_loop, n_steps);
    const size_t per = (size_t)h->W * h->N * 3;
    if (state_from_device(h, h->replay_poses + per * h->replay_frame, h->replay_vels + per * h->replay_frame, false)) return PMS_ERR_CUDA;
    h->a.steps_done += n_steps;
    return PMS_OK;
}

int pms_replay_step(pms_sim *h, double dt, int n_steps)
{
    HCHECK(h);
    int rc = replay_step_async(h, dt, n_steps);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return PMS_OK;
}

int pms_replay_frame(pms_sim *h, int set_frame, int *frame_out)
{
    HCHECK(h);
    if (!h->replay_frames) return fail(PMS_ERR_STATE, "no recorded trajectories: call pms_set_replay first");
    if (set_frame >= 0) {
        if (set_frame >= h->replay_frames) return fail(PMS_ERR_INVALID, "frame out of range");
        h->replay_frame = set_frame;
        const size_t per = (size_t)h->W * h->N * 3;
        if (state_from_device(h, h->replay_poses + per * set_frame, h->replay_vels + per * set_frame, false)) return PMS_ERR_CUDA;
        CK(cudaStreamSynchronize(h->stream));
    }
    if (frame_out) *frame_out = h->replay_frame;
    return PMS_OK;
}

int pms_replay_tick(pms_sim *h, int n_steps, const double *robot_pose, double robot_radius, uint32_t *coll_mask, uint32_t *det_mask,
                    double *det_vals, int32_t *frame_out)
{
    HCHECK(h);
    if (n_steps < 1) return fail(PMS_ERR_INVALID, "n_steps < 1");
    HsfmArgs &a = h->a;
    const int W = h->W;
    const size_t n = (size_t)W * h->N, nm = (size_t)W * h->words;
    const size_t o_coll = n * 16, o_det = o_coll + nm * 4, total = o_det + nm * 4;
    const size_t need = total > (size_t)W * 96 ? total : (size_t)W * 96;
    if (need > h->tick_bytes) {
        if (h->tick_dev) cudaFree(h->tick_dev);
        if (h->tick_host) cudaFreeHost(h->tick_host);
        h->tick_dev = h->tick_host = nullptr; h->tick_bytes = 0;
        CK(cudaMalloc((void **)&h->tick_dev, need));
        CK(cudaMallocHost((void **)&h->tick_host, need));
        h->tick_bytes = need;
    }
    if (!robot_pose) {
        a.robot_model = PMS_ROBOT_NONE;
    } else {
        if (!(robot_radius > 0.0)) return fail(PMS_ERR_INVALID, "robot radius must be > 0");
        a.robot_model = PMS_ROBOT_EXTERNAL; a.robot_radius = robot_radius;
        double *st = reinterpret_cast<double *>(h->tick_host);       // the previous tick was synchronised: the block is free
        std::memset(st, 0, sizeof(double) * W * 11 + sizeof(int) * W);
        std::memcpy(st, robot_pose, sizeof(double) * W * 3);
        CK(cudaMemcpyAsync(h->robot_block, st, sizeof(double) * W * 11 + sizeof(int) * W, cudaMemcpyHostToDevice, h->stream));
    }
    int rc = replay_step_async(h, a.dt > 0.0 ? a.dt : 0.01, n_steps);
    if (rc) return rc;
    unsigned char *d = h->tick_dev;
    CK(cudaMemcpyAsync(d, a.det_vals, n * 16, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(d + o_coll, a.coll_mask, nm * 4, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(d + o_det, a.det_mask, nm * 4, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(h->tick_host, d, total, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const unsigned char *q = h->tick_host;
    if (det_vals) std::memcpy(det_vals, q, n * 16);
    if (coll_mask) std::memcpy(coll_mask, q + o_coll, nm * 4);
    if (det_mask) std::memcpy(det_mask, q + o_det, nm * 4);
    if (frame_out) *frame_out = h->replay_frame;
    return PMS_OK;
}

// ---- sensors inside a fused launch ----
int pms_lidar_config(pms_sim *h, int n_beams, double max_dist, double resolution)
{
    HCHECK(h);
    if (n_beams <= 0) return fail(PMS_ERR_INVALID, "n_beams must be > 0");                          // _lidar.py:15
    if (!(max_dist > resolution && resolution > 0.0)) return fail(PMS_ERR_INVALID, "need max_dist > resolution > 0"); // :16
    pms_sim::Sched &sc = h->sched;
    std::vector<double> dirs;
    lidar_directions(n_beams, dirs);
    CK(cudaStreamSynchronize(h->stream));
    if (sc.lidar_dirs) cudaFree(sc.lidar_dirs);
    sc.lidar_dirs = nullptr;
    CK(cudaMalloc((void **)&sc.lidar_dirs, dirs.size() * sizeof(double)));
    CK(cudaMemcpy(sc.lidar_dirs, dirs.data(), dirs.size() * sizeof(double), cudaMemcpyHostToDevice));
    if (n_beams != sc.lidar_beams) {          // records are sized per beam count
        if (sc.lidar_hit) cudaFree(sc.lidar_hit);
        if (sc.lidar_pts) cudaFree(sc.lidar_pts);
        sc.lidar_hit = nullptr; sc.lidar_pts = nullptr; sc.lidar_cap = 0; sc.lidar_pts_cap = 0;
    }
    sc.lidar_beams = n_beams; sc.lidar_max = max_dist; sc.lidar_res = resolution;
    sc.lidar_samples = (int)std::ceil((max_dist - 0.0) / resolution);     // np.arange(0, max_dist, resolution), _lidar.py:112
    sc.n_fires[SENSOR_LIDAR] = 0;
    return PMS_OK;
}

int pms_sensor_schedule(pms_sim *h, int kind, int enabled, double period, double hold_time, int flags)
{
    HCHECK(h);
    if (kind < 0 || kind >= SENSOR_KINDS) return fail(PMS_ERR_INVALID, "unknown sensor kind");
    if (enabled && !(period >= 0.0)) return fail(PMS_ERR_INVALID, "period must be >= 0");
    if (enabled && !(hold_time >= 0.0)) return fail(PMS_ERR_INVALID, "hold_time must be >= 0");
    pms_sim::Sched &sc = h->sched;
    sc.enabled[kind] = enabled != 0; sc.period[kind] = period; sc.flags[kind] = flags;
    sc.hold_host[kind] = enabled ? hold_time : 0.0;
    sc.n_fires[kind] = 0;
    if (sc.hold_dev) {
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaMemcpy(sc.hold_dev + kind, &sc.hold_host[kind], sizeof(double), cudaMemcpyHostToDevice));
    }
    return PMS_OK;
}

int pms_get_sensor_fires(pms_sim *h, int kind, int32_t *steps, int cap, int *n_fires, double *hold_time)
{
    HCHECK(h);
    if (kind < 0 || kind >= SENSOR_KINDS) return fail(PMS_ERR_INVALID, "unknown sensor kind");
    pms_sim::Sched &sc = h->sched;
    if (!sc.enabled[kind]) return fail(PMS_ERR_STATE, "sensor kind is not scheduled");
    int n_dev = 0;
    double hold_dev = sc.hold_host[kind];
    if (sc.hold_dev) {
        CK(cudaMemcpyAsync(&n_dev, sc.count_dev + kind, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(&hold_dev, sc.hold_dev + kind, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        // the device scheduler is authoritative; the host mirror sized the records and must agree bit for bit
        if (n_dev != sc.n_fires[kind] || hold_dev != sc.hold_host[kind])
            return fail(PMS_ERR_STATE, "device scheduler and host mirror disagree");
    }
    const int n = sc.n_fires[kind];
    if (n_fires) *n_fires = n;
    if (hold_time) *hold_time = hold_dev;
    if (steps && n > 0) {
        if (cap < n) return fail(PMS_ERR_INVALID, "steps buffer too small");
        CK(cudaMemcpyAsync(steps, sc.tables + (size_t)(SENSOR_KINDS + kind) * sc.table_cap, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    return PMS_OK;
}

static int record_range(pms_sim *h, int kind, int first, int count)
{
    if (!h->sched.enabled[kind]) return fail(PMS_ERR_STATE, "sensor kind is not scheduled");
    if (first < 0 || count < 0 || first + count > h->sched.n_fires[kind]) return fail(PMS_ERR_INVALID, "record range outside the readings of the last launch");
    return 0;
}

int pms_get_detector_records(pms_sim *h, int first, int count, uint32_t *mask, double *values)
{
    HCHECK(h);
    int rc = record_range(h, SENSOR_DETECTOR, first, count);
    if (rc) return rc;
    const size_t per_m = (size_t)h->W * h->words, per_v = (size_t)h->W * h->N;
    if (values && !(h->sched.flags[SENSOR_DETECTOR] & PMS_SENSOR_VALUES)) return fail(PMS_ERR_STATE, "detector records were scheduled without PMS_SENSOR_VALUES");
    if (count == 0) return PMS_OK;
    if (mask) CK(cudaMemcpyAsync(mask, h->sched.det_mask + first * per_m, sizeof(uint32_t) * count * per_m, cudaMemcpyDeviceToHost, h->stream));
    if (values) CK(cudaMemcpyAsync(values, h->sched.det_vals + first * per_v, sizeof(double2) * count * per_v, cudaMemcpyDeviceToHost, h->stream));
    IO_DONE(h);
    return PMS_OK;
}

int pms_get_collision_records(pms_sim *h, int first, int count, uint32_t *mask)
{
    HCHECK(h);
    int rc = record_range(h, SENSOR_COLLISIONS, first, count);
    if (rc) return rc;
    const size_t per_m = (size_t)h->W * h->words;
    if (mask && count) CK(cudaMemcpyAsync(mask, h->sched.coll_mask + first * per_m, sizeof(uint32_t) * count * per_m, cudaMemcpyDeviceToHost, h->stream));
    IO_DONE(h);
    return PMS_OK;
}

int pms_get_lidar_records(pms_sim *h, int first, int count, int32_t *hit_index, double *points)
{
    HCHECK(h);
    int rc = record_range(h, SENSOR_LIDAR, first, count);
    if (rc) return rc;
    if (h->a.robot_model == ROBOT_NONE) return fail(PMS_ERR_STATE, "no robot configured");
    const size_t per = (size_t)h->W * h->sched.lidar_beams;
    if (count == 0) return PMS_OK;
    if (hit_index) CK(cudaMemcpyAsync(hit_index, h->sched.lidar_hit + first * per, sizeof(int32_t) * count * per, cudaMemcpyDeviceToHost, h->stream));
    if (points) CK(cudaMemcpyAsync(points, h->sched.lidar_pts + first * per, sizeof(double2) * count * per, cudaMemcpyDeviceToHost, h->stream));
    IO_DONE(h);
    return PMS_OK;
}

int pms_get_omni_records(pms_sim *h, int first, int count, double *min_dist)
{
    HCHECK(h);
    int rc = record_range(h, SENSOR_OMNI, first, count);
    if (rc) return rc;
    if (h->a.robot_model == ROBOT_NONE) return fail(PMS_ERR_STATE, "no robot configured");
    if (min_dist && count) CK(cudaMemcpyAsync(min_dist, h->sched.omni + (size_t)first * h->W, sizeof(double) * count * h->W, cudaMemcpyDeviceToHost, h->stream));
    IO_DONE(h);
    return PMS_OK;
}

int pms_get_launch_info(pms_sim *h, pms_launch_info *out)
{
    if (!h || !out) return fail(PMS_ERR_INVALID, "NULL argument");
    *out = h->info;
    return PMS_OK;
}

int pms_set_tuning(pms_sim *h, int j_split, int worlds_per_block, int n_chunks, int cluster_size)
{
    if (!h) return fail(PMS_ERR_INVALID, "handle is NULL");
    h->tune_S = j_split; h->tune_wpc = worlds_per_block; h->tune_chunks = n_chunks < 0 ? 0 : n_chunks;
    h->no_prune = n_chunks == -1;                     // -1: the large-crowd kernel visits every j-tile
    h->no_reorder = n_chunks < 0;                     // -1, -2: no sorted working copy
    h->force_large = cluster_size < 0;               // negative: force the large-crowd kernel (tests)
    h->tune_cluster = cluster_size == -100 ? 0 : (cluster_size < 0 ? -cluster_size : cluster_size);   // -100: forced, default kernel choice
    h->large.sorted_valid = false;
    return PMS_OK;
}

// FP32 FMA-pipe peak of the current device, measured with 8 independent FMA chains per thread.
__global__ void fma_peak_kernel(float *out, int iters)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f,
          a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float b = 0.999f, c = 1e-3f;
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

double pms_measure_fp32_tflops(int device)
{
    Guard gd(device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
    const int blocks = prop.multiProcessorCount * 4, threads = 512, iters = 4096;
    float *buf = nullptr;
    if (cudaMalloc((void **)&buf, sizeof(float) * blocks * threads) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        fma_peak_kernel<<<blocks, threads>>>(buf, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        if (rep >= 1 && ms > 0.f) best = fmax(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf);
    return best;
}

void *pms_host_alloc(uint64_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void pms_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---- ORCA ----
int pms_orca_create(int device, int n_worlds, int n_agents, const pms_orca_params *params, pms_sim **out)
{
    if (!out) return fail(PMS_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n_worlds < 1 || n_agents < 1) return fail(PMS_ERR_INVALID, "n_worlds and n_agents must be >= 1");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(PMS_ERR_CUDA, std::string("no CUDA device available (libpms_b200 has no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(PMS_ERR_INVALID, "device index out of range");
    Guard gd(device);
    pms_orca_params p;
    if (params) p = *params; else pms_default_orca_params(&p);
    if (p.max_neighbors < 0 || p.max_neighbors > ORCA_MAX_NEIGHBORS)
        return fail(PMS_ERR_UNSUPPORTED, "max_neighbors must be in [0, 16]");
    if (n_agents > 256) return fail(PMS_ERR_UNSUPPORTED, "ORCA worlds are limited to 256 agents");
    pms_sim *h = new pms_sim();
    h->kind = 1; h->device = device; h->W = n_worlds; h->N = n_agents;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return fail(PMS_ERR_CUDA, "stream"); }
    int rc = orca_alloc(h->orca, n_worlds, n_agents, p.time_step, p.neighbor_dist, p.max_neighbors, p.time_horizon,
                        p.radius, p.goal_reaching_threshold);
    if (rc) { pms_destroy(h); return fail(PMS_ERR_CUDA, "ORCA allocation failed"); }
    h->info = pms_launch_info{};
    *out = h;
    return PMS_OK;
}

} // extern "C"

extern "C" {
#define OCHECK(h)                                                                  \
    if (!(h)) return fail(PMS_ERR_INVALID, "handle is NULL");                      \
    if ((h)->kind != 1) return fail(PMS_ERR_STATE, "not an ORCA handle");          \
    Guard gd__((h)->device)

int pms_orca_set_agents(pms_sim *h, const double *positions, const double *velocities, const double *max_speeds,
                        const double *pref_speeds)
{
    OCHECK(h);
    if (!positions || !max_speeds) return fail(PMS_ERR_INVALID, "positions / max_speeds is NULL");
    const size_t n = (size_t)h->W * h->N;
    for (size_t k = 0; k < n; ++k) {   // _orca.py:56-62
        if (velocities) {
            const double sp = std::sqrt(velocities[2 * k] * velocities[2 * k] + velocities[2 * k + 1] * velocities[2 * k + 1]);
            if (!(sp <= max_speeds[k])) return fail(PMS_ERR_INVALID, "Initial speeds must be <= all corresponding max_speeds");
        }
        if (pref_speeds && !(pref_speeds[k] <= max_speeds[k]))
            return fail(PMS_ERR_INVALID, "All preferred_speeds must be <= corresponding max_speeds");
    }
    if (orca_set_agents(h->orca, positions, velocities, max_speeds, pref_speeds, h->stream)) return fail(PMS_ERR_CUDA, "orca_set_agents failed");
    return PMS_OK;
}

int pms_orca_set_waypoints_fixed(pms_sim *h, const double *table, int n_rows, const int32_t *start_index, double reach, int loop)
{
    OCHECK(h);
    if (!table || n_rows < 1) return fail(PMS_ERR_INVALID, "table is NULL or n_rows < 1");
    if (start_index)
        for (size_t k = 0; k < (size_t)h->W * h->N; ++k)
            if (start_index[k] < 0 || start_index[k] >= n_rows) return fail(PMS_ERR_INVALID, "start_index out of range");
    if (orca_set_waypoints(h->orca, table, n_rows, start_index, reach, loop, h->stream)) return fail(PMS_ERR_CUDA, "orca_set_waypoints failed");
    return PMS_OK;
}

int pms_orca_set_waypoints_current(pms_sim *h, const double *current)
{
    OCHECK(h);
    if (!current) return fail(PMS_ERR_INVALID, "current is NULL");
    if (orca_set_current_waypoints(h->orca, current, h->stream)) return fail(PMS_ERR_CUDA, "orca_set_current_waypoints failed");
    return PMS_OK;
}

int pms_orca_set_pref_refresh(pms_sim *h, int enabled)
{
    OCHECK(h);
    h->orca.refresh_pref = enabled != 0;
    return PMS_OK;
}

int pms_orca_step_async(pms_sim *h, int n_steps)
{
    OCHECK(h);
    if (n_steps < 0) return fail(PMS_ERR_INVALID, "n_steps < 0");
    if (!h->orca.agents_set) return fail(PMS_ERR_STATE, "agents are not set");
    if (!h->orca.waypoints_set) return fail(PMS_ERR_STATE, "Waypoint tracker is not initialized");
    if (n_steps == 0) return PMS_OK;
    long long launches = 0;
    if (orca_step(h->orca, n_steps, h->stream, &launches)) return fail(PMS_ERR_CUDA, std::string("orca launch: ") + cudaGetErrorString(cudaGetLastError()));
    h->info.kernel = 2;
    h->info.kernel_launches += launches;
    return PMS_OK;
}

int pms_orca_step(pms_sim *h, int n_steps)
{
    int rc = pms_orca_step_async(h, n_steps);
    if (rc) return rc;
    return pms_sync(h);
}

// device-side snapshot of the ORCA agents (positions, velocities, preferred velocities, tracker position)
static void orca_snapshot_regions(pms_sim *h, std::vector<std::pair<void *, size_t>> &r)
{
    const OrcaState &o = h->orca;
    const size_t n = (size_t)h->W * h->N;
    r = {{o.pos, n * sizeof(float2)}, {o.vel, n * sizeof(float2)}, {o.pref, n * sizeof(float2)}, {o.cur_wp, n * sizeof(double2)},
         {o.wp_index, n * sizeof(int)}, {o.heading, n * sizeof(double)}};
}
int pms_orca_snapshot_save(pms_sim *h)
{
    OCHECK(h);
    if (!h->orca.agents_set) return fail(PMS_ERR_STATE, "agents are not set");
    std::vector<std::pair<void *, size_t>> r;
    orca_snapshot_regions(h, r);
    for (void *p : h->snap_dst) if (p) cudaFree(p);
    h->snap_dst.assign(r.size(), nullptr);
    for (size_t k = 0; k < r.size(); ++k) {
        CK(cudaMalloc(&h->snap_dst[k], r[k].second ? r[k].second : 1));
        CK(cudaMemcpyAsync(h->snap_dst[k], r[k].first, r[k].second, cudaMemcpyDeviceToDevice, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    h->snap_src = r;
    h->has_snapshot = true;
    return PMS_OK;
}
int pms_orca_snapshot_restore(pms_sim *h)
{
    OCHECK(h);
    if (!h->has_snapshot) return fail(PMS_ERR_STATE, "no snapshot saved");
    for (size_t k = 0; k < h->snap_src.size(); ++k)
        CK(cudaMemcpyAsync(h->snap_src[k].first, h->snap_dst[k], h->snap_src[k].second, cudaMemcpyDeviceToDevice, h->stream));
    return PMS_OK;
}

// one tick of the drop-in ORCAPedestriansModel: n_steps doStep + the agents back, one launch chain, ONE pinned transfer, one
// host synchronisation
int pms_orca_tick(pms_sim *h, int n_steps, double *poses, double *velocities)
{
    OCHECK(h);
    if (!poses || !velocities) return fail(PMS_ERR_INVALID, "null output");
    const size_t n = (size_t)h->W * h->N, need = n * 48;
    if (need > h->tick_bytes) {
        CK(cudaStreamSynchronize(h->stream));
        if (h->tick_dev) cudaFree(h->tick_dev);
        if (h->tick_host) cudaFreeHost(h->tick_host);
        h->tick_dev = h->tick_host = nullptr; h->tick_bytes = 0;
        CK(cudaMalloc((void **)&h->tick_dev, need));
        CK(cudaMallocHost((void **)&h->tick_host, need));
        h->tick_bytes = need;
    }
    int rc = pms_orca_step_async(h, n_steps);
    if (rc) return rc;
    if (orca_get_agents_async(h->orca, reinterpret_cast<double *>(h->tick_host), h->stream)) return fail(PMS_ERR_CUDA, "orca_get_agents failed");
    CK(cudaStreamSynchronize(h->stream));
    std::memcpy(poses, h->tick_host, n * 24);
    std::memcpy(velocities, h->tick_host + n * 24, n * 24);
    return PMS_OK;
}

int pms_orca_get_agents(pms_sim *h, double *poses, double *velocities)
{
    OCHECK(h);
    if (orca_get_agents(h->orca, poses, velocities, nullptr, h->stream)) return fail(PMS_ERR_CUDA, "orca_get_agents failed");
    return PMS_OK;
}
}
