// pms_common.cuh -- shared types and device math for libpms_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <math.h>

#define PMS_PI 3.14159265358979323846
#define PMS_TWO_PI 6.28318530717958647692
#define PMS_PED_RADIUS_CONST 0.3  // core/_constants.py:2 -- the collision list always uses this radius

namespace pms {

// Model constants pre-converted on the host to the arithmetic type that consumes them, so the SASS
// reads them as constant-bank operands instead of pinning registers for the whole step loop.
template <typename T> struct PairK {
    T rij;      // radius_i + radius_j
    T cexp;     // fp32: log2(e)/B ; fp64: B itself (the division of the reference is kept)
    T A, k1, k2;
};
template <typename T> struct PedConsts {
    T m, tau, ko, kd, kl, alpha, I, dt, inv_m, inv_tau, inv_alpha, one_alpha, m_over_tau, reach;
};

// ---------------------------------------------------------------------------------------------
// Sensors inside a fused launch.  The reference's hold-time scheduler (Simulation._update_sensors_state,
// core/_simulation.py:162-185: hold += dt; fire and reset when hold >= period) does not depend on the state, so
// sched_kernel (sensor_fires.cuh) replays it in fp64 for the n_steps of the launch before the step kernel runs and leaves,
// per sensor kind, fire[t] = slot of the reading taken after launch step t, or -1.  The step kernels write the readings of a
// firing step to record s of the kind's ring; the collision list (not a sensor in the reference: WorldState carries
// it every step, core/_simulation.py:141-143) is the pseudo-kind SENSOR_COLLISIONS.
// ---------------------------------------------------------------------------------------------
enum { SENSOR_DETECTOR = 0, SENSOR_LIDAR = 1, SENSOR_OMNI = 2, SENSOR_COLLISIONS = 3, SENSOR_KINDS = 4 };

struct LidarNoise {          // LidarSensorNoise (sensors/_lidar.py:52-87); l** = Cholesky factor of point_cov
    double drop_prob, mean_x, mean_y, l00, l10, l11;
    unsigned long long seed;
    int enabled;
};
struct OmniNoise {           // OmniObstacleDetectorNoise (sensors/_omni_obstacle_detector.py:25-37,64-69)
    double p_miss, mu, sigma;
    unsigned long long seed;
    int enabled;
};

struct SensorPlan {
    int on;                          // any kind scheduled for this launch
    int ped_on;                      // detector or collision-list records (written by the pedestrian warps)
    int robot_on;                    // lidar or omni records (written by the warp that integrates the robot)
    const int *fire[SENSOR_KINDS];   // (n_steps) record slot per launch step, -1 = no reading; NULL = kind not scheduled
    // detector records: mask (F, W, words); values (F, W, N) or NULL (masks only)
    uint32_t *det_mask;
    double2 *det_vals;
    // collision-list records (F, W, words)
    uint32_t *coll_mask;
    // lidar records: first occupied sample index or -1 (F, W, B); absolute sample point (F, W, B)
    int lidar_beams, lidar_samples;
    double lidar_res;
    const double *lidar_dirs;        // (B, 2) cos, sin of the beam angles, computed on the host with the C library
    int32_t *lidar_hit;
    double2 *lidar_pts;
    LidarNoise lidar_noise;
    // omni records (F, W)
    double *omni_dist;
    OmniNoise omni_noise;
};

// ---------------------------------------------------------------------------------------------
// Kernel argument block (passed by value as a __grid_constant__).
// ---------------------------------------------------------------------------------------------
struct HsfmArgs {
    int W, N, Np, wpc;          // worlds, peds/world, padded peds (multiple of 32), worlds per CTA
    int n_steps;
    int n_chunks, chunk_steps;  // dynamic (chunk, group) scheduling when n_chunks > 1
    int *work_counter, *group_done;
    long long steps_done;       // cumulative steps before this launch (non-finite bookkeeping)
    // pedestrian state, structure-of-arrays, world-major (element g = w*N + i)
    void *pos;                  // FP32: float4 {x,y,vx,vy}   | fp64 state: double2 {x,y}
    void *vel;                  // FP32: unused               | fp64 state: double2 {vx,vy}
    void *th;                   // {theta, omega}  float2 | double2
    // FP32 mode only: low parts of the state carried as unevaluated sums hi + lo (compensated Euler updates in
    // hsfm_warp.cuh; the other kernels write high parts only and the API zeroes these after their launches)
    void *pos_lo;               // float4 {x,y,vx,vy} low parts
    void *th_lo;                // float2 {theta, omega} low parts
    int comp;                   // 1: FP32 kernels carry the compensated state (PMS_FP32), 0: plain fp32 state (PMS_FP32_PLAIN)
    void *wp;                   // current waypoint float2 | double2
    void *vmag;                 // float | double
    int *wp_index;
    // waypoint tracker
    int tracker_kind, loop, n_table, q_cap;
    double reach;
    const void *table;          // (W*N*n_table) float2 | double2
    const void *queue;          // (W*N*q_cap) float2 | double2 ring
    int *q_head, *q_count;
    int *events, *ev_count;     // (ev_cap,3) step, world, ped
    int ev_cap;
    // robot (always fp64)
    int robot_model, robot_visible;
    double robot_radius, wheel_base;
    double *robot_pose, *robot_vel, *robot_ctrl, *robot_rear;
    const double *controls;     // (n_controls, W, 2)
    int n_controls;
    int *robot_world_collision;
    // map
    int map_kind, map_count, map_stride;
    const double *map_data;
    // detector + collision list
    int det_enabled, det_type;
    double det_max_dist, det_half_fov;
    uint32_t *det_mask, *coll_mask;
    double2 *det_vals;
    unsigned long long *det_count, *coll_count;
    int *nonfinite;
    const double *accel_noise;  // (W,N,2) added to the body-frame acceleration dv_b, or NULL (_hsfm.py:288-289)
    // device-side noise (counter-based Philox4x32-10, keyed by seed / pedestrian / absolute step: independent of the launch
    // configuration): N(0, accel_noise_std) drawn afresh every step like np.random.normal in _hsfm.py:289 (statistical parity)
    double accel_noise_std;
    unsigned long long noise_seed;
    // model constants (fp64 masters; kernels down-convert once)
    double dt;
    double tau, A, B, k1, k2, ko, kd, k_lambda, alpha, mass, ped_radius, hsfm_robot_radius;
    // derived blocks, filled by fill_consts() on the host before every launch
    PairK<float> kpp_f, kpr_f;       // ped-ped, ped-robot
    PairK<double> kpp_d, kpr_d;
    PedConsts<float> kc_f;
    PedConsts<double> kc_d;
    double sense_lim2;               // squared cheap-reject radius of the sensor pass
    // hsfm_warp.cuh: exponent offset r_ij*log2(e)/B rounded to fp32, and A rescaled by the rounding residue
    float warp_rc, warp_A;
    float warp_cut2;                 // squared distance beyond which ex2.approx.ftz of the pair exponent is exactly 0
    // fp32 pre-filter of the collision list / detector: squared radii 1e-5 inside and outside the exact thresholds and
    // cos(fov/2); only pedestrians between the two (or detected on the last step) reach the fp64 predicate
    float sn_coll_lo2, sn_coll_hi2, sn_det_lo2, sn_det_hi2, sn_any_hi2, sn_chf;
    SensorPlan sp;                   // sensors inside the launch (all zero = off)
};

template <typename T> __host__ __device__ inline void fill_pair(PairK<T> &k, const HsfmArgs &a, double radius_j)
{
    k.rij = (T)(a.ped_radius + radius_j);
    k.cexp = sizeof(T) == 4 ? (T)(1.4426950408889634 / a.B) : (T)a.B;
    k.A = (T)a.A; k.k1 = (T)a.k1; k.k2 = (T)a.k2;
}
template <typename T> __host__ __device__ inline void fill_ped(PedConsts<T> &c, const HsfmArgs &a)
{
    c.m = (T)a.mass; c.tau = (T)a.tau; c.ko = (T)a.ko; c.kd = (T)a.kd; c.kl = (T)a.k_lambda;
    c.alpha = (T)a.alpha; c.I = (T)(0.5 * (a.ped_radius * a.ped_radius)); c.dt = (T)a.dt;
    c.inv_m = (T)(1.0 / a.mass); c.inv_tau = (T)(1.0 / a.tau); c.inv_alpha = (T)(1.0 / a.alpha);
    c.one_alpha = (T)(1.0 + a.alpha); c.m_over_tau = (T)(a.mass / a.tau); c.reach = (T)a.reach;
}
inline void fill_consts(HsfmArgs &a)
{
    fill_pair(a.kpp_f, a, a.ped_radius); fill_pair(a.kpr_f, a, a.hsfm_robot_radius);
    fill_pair(a.kpp_d, a, a.ped_radius); fill_pair(a.kpr_d, a, a.hsfm_robot_radius);
    fill_ped(a.kc_f, a); fill_ped(a.kc_d, a);
    const double coll_r = a.robot_radius + PMS_PED_RADIUS_CONST;
    const double lim = (a.det_enabled ? (a.det_max_dist > coll_r ? a.det_max_dist : coll_r) : coll_r) * 1.001 + 1e-3;
    a.sense_lim2 = lim * lim;
    {
        const double eps = 1e-5, lo = (1.0 - eps) * (1.0 - eps), hi = (1.0 + eps) * (1.0 + eps);
        a.sn_coll_lo2 = (float)(coll_r * coll_r * lo); a.sn_coll_hi2 = (float)(coll_r * coll_r * hi);
        const double md = a.det_enabled && a.det_max_dist > 0.0 ? a.det_max_dist : 0.0;
        a.sn_det_lo2 = (float)(md * md * lo); a.sn_det_hi2 = a.det_enabled ? (float)(md * md * hi) : -1.0f;
        a.sn_any_hi2 = a.sn_coll_hi2 > a.sn_det_hi2 ? a.sn_coll_hi2 : a.sn_det_hi2;
        // |diff| <= fov/2  <=>  cos(diff) >= cos(fov/2) for 0 <= fov/2 < pi; always true beyond, never for fov < 0
        a.sn_chf = a.det_half_fov < 0.0 ? 2.0f : (a.det_half_fov >= PMS_PI ? -2.0f : (float)cos(a.det_half_fov));
    }
    const double rc = (a.ped_radius + a.ped_radius) * (1.4426950408889634 / a.B);
    a.warp_rc = (float)rc;
    a.warp_A = (float)(a.A * exp2(rc - (double)a.warp_rc));
    {   // (r_ij - d) log2(e)/B < -127.5 (ex2.approx.ftz flushes results below 2^-126 to exactly 0), 1e-3 relative margin;
        // r_ij: the larger of the pedestrian-pedestrian and pedestrian-robot radius sums
        const double rmax = a.ped_radius + (a.ped_radius > a.hsfm_robot_radius ? a.ped_radius : a.hsfm_robot_radius);
        const double cut = (rmax + 127.5 * a.B / 1.4426950408889634) * (1.0 + 1e-3);
        a.warp_cut2 = a.B > 0.0 ? (float)(cut * cut) : 3.0e38f;
    }
}
__device__ __forceinline__ const PairK<float> &kpp_of(const HsfmArgs &a, float) { return a.kpp_f; }
__device__ __forceinline__ const PairK<double> &kpp_of(const HsfmArgs &a, double) { return a.kpp_d; }
__device__ __forceinline__ const PairK<float> &kpr_of(const HsfmArgs &a, float) { return a.kpr_f; }
__device__ __forceinline__ const PairK<double> &kpr_of(const HsfmArgs &a, double) { return a.kpr_d; }
__device__ __forceinline__ const PedConsts<float> &kc_of(const HsfmArgs &a, float) { return a.kc_f; }
__device__ __forceinline__ const PedConsts<double> &kc_of(const HsfmArgs &a, double) { return a.kc_d; }

enum { ROBOT_NONE = 0, ROBOT_EXTERNAL = 1, ROBOT_UNICYCLE = 2, ROBOT_BICYCLE = 3 };
enum { TRACKER_FIXED = 0, TRACKER_QUEUE = 1 };
enum { MAP_EMPTY = 0, MAP_CIRCLES = 1, MAP_AABB = 2 };
enum { DET_POLAR = 0, DET_RELATIVE_ROTATED = 1, DET_RELATIVE = 2, DET_ABSOLUTE = 3 };

// ---------------------------------------------------------------------------------------------
// Overloaded scalar math (float / double) so the per-pedestrian code is written once.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float m_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double m_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float m_atan2(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double m_atan2(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ void m_sincos(float a, float *s, float *c) { sincosf(a, s, c); }
__device__ __forceinline__ void m_sincos(double a, double *s, double *c) { sincos(a, s, c); }
__device__ __forceinline__ float m_fmod(float a, float b) { return fmodf(a, b); }
__device__ __forceinline__ double m_fmod(double a, double b) { return fmod(a, b); }
__device__ __forceinline__ float m_exp(float a) { return expf(a); }
__device__ __forceinline__ double m_exp(double a) { return exp(a); }
__device__ __forceinline__ float m_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double m_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ bool m_finite(float a) { return isfinite(a); }
__device__ __forceinline__ bool m_finite(double a) { return isfinite(a); }

// ---------------------------------------------------------------------------------------------
// Counter-based random numbers: Philox4x32-10 (Salmon et al., SC'11), one 128-bit block per (seed, stream, id, step).
// ---------------------------------------------------------------------------------------------
enum { NOISE_STREAM_ACCEL = 0, NOISE_STREAM_DETECTOR = 1, NOISE_STREAM_LIDAR = 2, NOISE_STREAM_OMNI = 3 };

__host__ __device__ inline uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = 0xD2511F53ull * c.x, p1 = 0xCD9E8D57ull * c.z;
        c = make_uint4((unsigned)(p1 >> 32) ^ c.y ^ k.x, (unsigned)p1, (unsigned)(p0 >> 32) ^ c.w ^ k.y, (unsigned)p0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
__host__ __device__ inline uint4 noise_block(unsigned long long seed, int stream, unsigned long long id, unsigned long long step)
{
    return philox4x32_10(make_uint4((unsigned)id, (unsigned)(id >> 32), (unsigned)step, (unsigned)(step >> 32) ^ ((unsigned)stream << 28)),
                         make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
}
// uniform in (0, 1): 32 random bits, centred in their cell
__host__ __device__ inline double noise_uniform(unsigned bits) { return ((double)bits + 0.5) * (1.0 / 4294967296.0); }
// two independent standard normals from two 32-bit words (Box-Muller, fp64: accurate far into the tails)
static __device__ __noinline__ double2 noise_normal2(unsigned b0, unsigned b1)
{
    const double r = sqrt(-2.0 * log(noise_uniform(b0)));
    double s, c;
    sincospi(2.0 * noise_uniform(b1), &s, &c);
    return make_double2(r * c, r * s);
}
// body-frame acceleration noise of pedestrian g at absolute step `step` (_hsfm.py:288-289)
static __device__ __noinline__ double2 accel_noise_draw(const HsfmArgs &a, size_t g, long long step)
{
    const uint4 b = noise_block(a.noise_seed, NOISE_STREAM_ACCEL, g, (unsigned long long)step);
    const double2 z = noise_normal2(b.x, b.y);
    return make_double2(a.accel_noise_std * z.x, a.accel_noise_std * z.y);
}

// Python / numpy float modulo (sign of the divisor): _hsfm.py:98,189, util/_util.py:6.
template <typename T>
__device__ __forceinline__ T py_mod(T a, T b)
{
    T m = m_fmod(a, b);
    if (m != T(0)) {
        if ((b < T(0)) != (m < T(0))) m += b;
    } else {
        m = copysign(T(0), b);
    }
    return m;
}
template <typename T>
__device__ __forceinline__ T wrap_angle(T a)
{
    return py_mod(a + T(PMS_PI), T(PMS_TWO_PI)) - T(PMS_PI);
}

template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };
template <typename T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<double> { using type = double4; };

// ---------------------------------------------------------------------------------------------
// State I/O: FP32 state is one float4 {x,y,vx,vy} per pedestrian (one coalesced 16-byte load per
// lane); fp64 state is two double2.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_pv(const HsfmArgs &a, size_t g, float &x, float &y, float &vx, float &vy)
{
    float4 p = __ldcg(reinterpret_cast<const float4 *>(a.pos) + g);
    x = p.x; y = p.y; vx = p.z; vy = p.w;
}
__device__ __forceinline__ void load_pv(const HsfmArgs &a, size_t g, double &x, double &y, double &vx, double &vy)
{
    double2 p = __ldcg(reinterpret_cast<const double2 *>(a.pos) + g);
    double2 v = __ldcg(reinterpret_cast<const double2 *>(a.vel) + g);
    x = p.x; y = p.y; vx = v.x; vy = v.y;
}
__device__ __forceinline__ void store_pv(const HsfmArgs &a, size_t g, float x, float y, float vx, float vy)
{
    reinterpret_cast<float4 *>(a.pos)[g] = make_float4(x, y, vx, vy);
}
__device__ __forceinline__ void store_pv(const HsfmArgs &a, size_t g, double x, double y, double vx, double vy)
{
    reinterpret_cast<double2 *>(a.pos)[g] = make_double2(x, y);
    reinterpret_cast<double2 *>(a.vel)[g] = make_double2(vx, vy);
}

// ---------------------------------------------------------------------------------------------
// Pair force  F(i <- j)  (pedestrians/_hsfm.py:113-134).
// ---------------------------------------------------------------------------------------------
// Single-instruction MUFU forms (the library rsqrtf/exp2f add denormal fix-ups we do not need).
__device__ __forceinline__ float fast_rsqrt(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_ex2(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// fp32 pair force, split in two parts so that the common case (no body contact, g = 0) costs
// 14 FMA-pipe + 2 MUFU instructions:
//   pair_far     : exponential repulsion  A exp((r_ij - d)/B) n   (always)
//   pair_contact : k_1 g n + k_2 g dv_t t                          (only when some lane has g > 0)
// d = d2 * rsqrt(d2) (MUFU.RSQ, ~2 ulp).  -DPMS_REFINE_SQRT adds one Newton step; measured to make no
// difference to parity with the reference because fp32 state rounding dominates.
// `self` zeroes the j == i term exactly (inv = 0 => every product below is 0).
__device__ __forceinline__ void pair_far(const PairK<float> &k, float dx, float dy, bool self,
                                         float &inv, float &s, float &fx, float &fy)
{
    const float d2 = fmaf(dx, dx, dy * dy);
    inv = fast_rsqrt(d2);
    inv = self ? 0.0f : inv;
    float d = d2 * inv;
#ifdef PMS_REFINE_SQRT
    const float err = fmaf(-d, d, d2);
    d = fmaf(err, 0.5f * inv, d);
#endif
    s = k.rij - d;
    const float e = fast_ex2(s * k.cexp);
    const float a = e * (k.A * inv);             // A e / d : multiplies (dx,dy) directly
    fx = fmaf(a, dx, fx);
    fy = fmaf(a, dy, fy);
}
__device__ __forceinline__ void pair_contact(const PairK<float> &k, float dx, float dy, float dvx, float dvy,
                                             float inv, float s, float &fx, float &fy)
{
    const float g = fmaxf(s, 0.0f);
    const float w = fmaf(dvy, dx, -dvx * dy);    // d * (v_j - v_i) . t
    const float gi = g * inv;
    const float a = k.k1 * gi;                   // k_1 g / d
    const float b = k.k2 * gi * inv * w;         // k_2 g dv_t / d
    fx = fmaf(a, dx, fx); fx = fmaf(-b, dy, fx);
    fy = fmaf(a, dy, fy); fy = fmaf(b, dx, fy);
}
__device__ __forceinline__ void pair_accum(const PairK<float> &k, float dx, float dy, float dvx, float dvy,
                                           bool self, float &fx, float &fy)
{
    float inv, s;
    pair_far(k, dx, dy, self, inv, s, fx, fy);
    pair_contact(k, dx, dy, dvx, dvy, inv, s, fx, fy);
}

// fp64: same operation order as the reference.
__device__ __forceinline__ void pair_accum(const PairK<double> &k, double dx, double dy, double dvx, double dvy,
                                           bool self, double &fx, double &fy)
{
    if (self) return;
    const double d = sqrt(dx * dx + dy * dy);
    const double nx = dx / d, ny = dy / d;
    const double tx = -ny, ty = nx;
    const double dvt = dvx * tx + dvy * ty;
    double g = k.rij - d;
    if (!(g > 0.0)) g = 0.0;
    const double cn = k.A * exp((k.rij - d) / k.cexp) + k.k1 * g;
    double px = cn * nx, py = cn * ny;
    const double ct = k.k2 * g * dvt;
    px += ct * tx;
    py += ct * ty;
    fx += px;
    fy += py;
}

// ---------------------------------------------------------------------------------------------
// World map predicates (fp64), shared by the robot veto, lidar and omni kernels.
// world_map/_circles_world.py:29-50, _aabb_world.py:98-113,135-156, util/_geometry.py:4-42.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool map_occupied(int kind, int count, const double *m, double x, double y)
{
    if (kind == MAP_CIRCLES) {
        for (int c = 0; c < count; ++c) {
            const double dx = x - m[3 * c], dy = y - m[3 * c + 1];
            if (sqrt(dx * dx + dy * dy) - m[3 * c + 2] <= 0.0) return true;
        }
    } else if (kind == MAP_AABB) {
        for (int c = 0; c < count; ++c) {
            const double tlx = m[4 * c], tly = m[4 * c + 1], bw = m[4 * c + 2], bh = m[4 * c + 3];
            if (x <= tlx && x >= (tlx - bh) && y >= tly && y <= (tly + bw)) return true;
        }
    }
    return false;
}

__device__ __forceinline__ double seg_dist(double px, double py, double x1, double y1, double x2, double y2)
{
    const double ex = x2 - x1, ey = y2 - y1;
    const double ox = px - x1, oy = py - y1;
    double t = (ox * ex + oy * ey) / (ex * ex + ey * ey);
    t = t < 0.0 ? 0.0 : t;   // NaN (degenerate segment) falls through both clamps like np.clip
    t = t > 1.0 ? 1.0 : t;
    const double qx = x1 + t * ex, qy = y1 + t * ey;
    const double dx = px - qx, dy = py - qy;
    return sqrt(dx * dx + dy * dy);
}

__device__ __forceinline__ double map_closest(int kind, int count, const double *m, double x, double y)
{
    double best = CUDART_INF;
    if (kind == MAP_CIRCLES) {
        for (int c = 0; c < count; ++c) {
            const double dx = x - m[3 * c], dy = y - m[3 * c + 1];
            const double d = sqrt(dx * dx + dy * dy) - m[3 * c + 2];
            if (d < best || d != d) best = d;
        }
    } else if (kind == MAP_AABB) {
        for (int c = 0; c < count; ++c) {
            const double tlx = m[4 * c], tly = m[4 * c + 1], bw = m[4 * c + 2], bh = m[4 * c + 3];
            double d;
            d = seg_dist(x, y, tlx, tly, tlx, tly + bw);            if (d < best || d != d) best = d;
            d = seg_dist(x, y, tlx - bh, tly, tlx - bh, tly + bw);  if (d < best || d != d) best = d;
            d = seg_dist(x, y, tlx, tly, tlx - bh, tly);            if (d < best || d != d) best = d;
            d = seg_dist(x, y, tlx, tly + bw, tlx - bh, tly + bw);  if (d < best || d != d) best = d;
        }
    }
    return best;
}

// core/_world_map.py:35-62 with eps_margin = 0.05
__device__ __forceinline__ bool map_is_collision(int kind, int count, const double *m, double x, double y, double radius)
{
    if (kind == MAP_EMPTY || count == 0) return false;
    if (kind == MAP_CIRCLES) {          // is_occupied and closest-distance share |p - c| - r: one pass, the same values
        bool occ = false;
        double best = CUDART_INF;
        for (int c = 0; c < count; ++c) {
            const double dx = x - m[3 * c], dy = y - m[3 * c + 1];
            const double d = sqrt(dx * dx + dy * dy) - m[3 * c + 2];
            occ = occ || d <= 0.0;
            if (d < best || d != d) best = d;
        }
        return occ || (best - radius + 0.05 <= 0.0);
    }
    const bool occ = map_occupied(kind, count, m, x, y);
    const double dist = map_closest(kind, count, m, x, y) - radius + 0.05;
    return occ || (dist <= 0.0);
}

// ---- the same predicates evaluated by a whole warp: obstacle c on lane c & 31, results reduced with shuffles ----
// (the per-obstacle arithmetic is that of map_closest / map_is_collision; the reduction -- min, any, NaN wins -- returns what
// the sequential loops return)
__device__ __forceinline__ double map_closest_warp(int kind, int count, const double *m, double x, double y, int lane, bool *occupied = nullptr)
{
    double best = CUDART_INF;
    bool nan_seen = false, occ = false;
    if (kind == MAP_CIRCLES) {
        for (int c = lane; c < count; c += 32) {
            const double dx = x - m[3 * c], dy = y - m[3 * c + 1];
            const double d = sqrt(dx * dx + dy * dy) - m[3 * c + 2];
            occ = occ || d <= 0.0;
            nan_seen = nan_seen || d != d;
            if (d < best) best = d;
        }
    } else if (kind == MAP_AABB) {
        for (int c = lane; c < count; c += 32) {
            const double tlx = m[4 * c], tly = m[4 * c + 1], bw = m[4 * c + 2], bh = m[4 * c + 3];
            occ = occ || (x <= tlx && x >= (tlx - bh) && y >= tly && y <= (tly + bw));
            double d;
            d = seg_dist(x, y, tlx, tly, tlx, tly + bw);            nan_seen = nan_seen || d != d; if (d < best) best = d;
            d = seg_dist(x, y, tlx - bh, tly, tlx - bh, tly + bw);  nan_seen = nan_seen || d != d; if (d < best) best = d;
            d = seg_dist(x, y, tlx, tly, tlx - bh, tly);            nan_seen = nan_seen || d != d; if (d < best) best = d;
            d = seg_dist(x, y, tlx, tly + bw, tlx - bh, tly + bw);  nan_seen = nan_seen || d != d; if (d < best) best = d;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, best, o);
        if (other < best) best = other;
    }
    if (occupied) *occupied = __any_sync(0xffffffffu, occ);
    // the sequential rule `if (d < best || d != d) best = d` keeps a NaN once one was seen (nothing compares below it): any NaN => NaN
    if (__any_sync(0xffffffffu, nan_seen)) best = CUDART_NAN;
    return best;
}
__device__ __forceinline__ bool map_is_collision_warp(int kind, int count, const double *m, double x, double y, double radius, int lane)
{
    if (kind == MAP_EMPTY || count == 0) return false;
    bool occ = false;
    const double dist = map_closest_warp(kind, count, m, x, y, lane, &occ) - radius + 0.05;
    return occ || (dist <= 0.0);
}

} // namespace pms
