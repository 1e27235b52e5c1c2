// sensors.cuh -- stand-alone sensor kernels on the current state: collision list + pedestrian
// detector (no stepping), lidar ray march against the map, omni obstacle distance.
#pragma once
#include "hsfm_small.cuh"

namespace pms {

// Collision list + detector on the stored state (Simulation.__init__ sensor pass, core/_simulation.py:37).
// Grid (world, block of 256 pedestrians), one thread per (padded) pedestrian.
template <typename ST>
__global__ void sense_kernel(const __grid_constant__ HsfmArgs a)
{
    const int w = blockIdx.x, i = blockIdx.y * blockDim.x + threadIdx.x;
    const int N = a.N, words = (N + 31) >> 5;
    RobotSlot rb{a.robot_pose[3 * w], a.robot_pose[3 * w + 1], a.robot_pose[3 * w + 2], 0.0, 0.0};
    SenseOut so{false, false, 0.0, 0.0};
    const size_t g = (size_t)w * N + i;
    if (i < N) {
        ST x, y, vx, vy;
        load_pv(a, g, x, y, vx, vy);
        if constexpr (sizeof(ST) == 4) {          // FP32 state = hi + lo (lo is zero unless the compensated kernel ran)
            const float4 l = reinterpret_cast<const float4 *>(a.pos_lo)[g];
            so = sense_ped<double>(a, rb, (double)x + (double)l.x, (double)y + (double)l.y, true);
        } else {
            so = sense_ped<ST>(a, rb, x, y, true);
        }
        if (a.det_enabled) a.det_vals[g] = make_double2(so.v0, so.v1);
    }
    const unsigned dm = __ballot_sync(0xffffffffu, so.det);
    const unsigned cm = __ballot_sync(0xffffffffu, so.coll);
    if ((i & 31) == 0 && (i >> 5) < words) {
        a.det_mask[(size_t)w * words + (i >> 5)] = dm;
        a.coll_mask[(size_t)w * words + (i >> 5)] = cm;
    }
}

// PedestrianDetectorNoise on the device (sensors/_pedestrian_detector.py:104-112 + _convert_reading :114-134): every detected
// pedestrian is dropped with probability misdetection_prob, the survivors get N(mu, sigma) added to distance and angle in
// the polar frame and are then converted to the configured return type.  `vals_in` holds POLAR readings (the step kernels
// are switched to polar while the noise is on).  Philox block per (world-pedestrian, absolute step): deterministic.
struct DetNoise {
    double p_miss, d_mu, d_sigma, a_mu, a_sigma;
    unsigned long long seed;
    int enabled, return_type;
};
__global__ void detector_noise_kernel(int W, int N, int words, const uint32_t *mask_in, const double2 *vals_in,
                                      uint32_t *mask_out, double2 *vals_out, const double *robot_pose, DetNoise nz,
                                      long long step)
{
    const int w = blockIdx.x, i = blockIdx.y * blockDim.x + threadIdx.x;      // one thread per padded pedestrian
    const size_t g = (size_t)w * N + i;
    bool det = i < N && ((mask_in[(size_t)w * words + (i >> 5)] >> (i & 31)) & 1u);
    double2 v = make_double2(0.0, 0.0);
    if (det) {
        const uint4 b = noise_block(nz.seed, NOISE_STREAM_DETECTOR, g, (unsigned long long)step);
        if (noise_uniform(b.x) < nz.p_miss) det = false;                 // np.random.binomial(1, 1 - p) == 0
        else {
            const double2 z = noise_normal2(b.y, b.z);
            const double2 p = vals_in[g];
            const double distance = p.x + (nz.d_mu + nz.d_sigma * z.x), angle = p.y + (nz.a_mu + nz.a_sigma * z.y);
            if (nz.return_type == DET_POLAR) v = make_double2(distance, angle);
            else {
                double sa, ca;
                sincos(angle, &sa, &ca);
                const double xrr = distance * ca, yrr = distance * sa;
                if (nz.return_type == DET_RELATIVE_ROTATED) v = make_double2(xrr, yrr);
                else {
                    double st, ct;
                    sincos(robot_pose[3 * w + 2], &st, &ct);
                    const double xr = xrr * ct - yrr * st, yr = xrr * st + yrr * ct;
                    v = nz.return_type == DET_RELATIVE ? make_double2(xr, yr)
                                                       : make_double2(xr + robot_pose[3 * w], yr + robot_pose[3 * w + 1]);
                }
            }
        }
    }
    if (i < N) vals_out[g] = v;
    const unsigned dm = __ballot_sync(0xffffffffu, det);
    if ((i & 31) == 0 && (i >> 5) < words) mask_out[(size_t)w * words + (i >> 5)] = dm;
}

// LidarSensorNoise.noisify_reading (sensors/_lidar.py:79-87): every hit point is dropped with probability drop_prob (its
// hit index becomes -1), the survivors get mean + L z added, L = Cholesky factor of point_cov, z ~ N(0, I).
struct LidarNoise {
    double drop_prob, mean_x, mean_y, l00, l10, l11;
    unsigned long long seed;
    int enabled;
};
__global__ void lidar_noise_kernel(size_t nb, int32_t *hit, double2 *points, LidarNoise nz, long long step)
{
    const size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k >= nb || hit[k] < 0) return;
    const uint4 b = noise_block(nz.seed, NOISE_STREAM_LIDAR, k, (unsigned long long)step);
    if (noise_uniform(b.x) < nz.drop_prob) { hit[k] = -1; points[k] = make_double2(0.0, 0.0); return; }
    const double2 z = noise_normal2(b.y, b.z);
    double2 p = points[k];
    p.x += nz.mean_x + nz.l00 * z.x;
    p.y += nz.mean_y + (nz.l10 * z.x + nz.l11 * z.y);
    points[k] = p;
}

// OmniObstacleDetector._noisify_reading (sensors/_omni_obstacle_detector.py:64-69): inf with probability
// misdetection_prob, else distance + N(mu, sigma).
struct OmniNoise {
    double p_miss, mu, sigma;
    unsigned long long seed;
    int enabled;
};
__global__ void omni_noise_kernel(int W, double *dist, OmniNoise nz, long long step)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    const uint4 b = noise_block(nz.seed, NOISE_STREAM_OMNI, (unsigned long long)w, (unsigned long long)step);
    if (noise_uniform(b.x) < nz.p_miss) { dist[w] = CUDART_INF; return; }
    dist[w] += nz.mu + nz.sigma * noise_normal2(b.y, b.z).x;
}

struct LidarArgs {
    int W, n_beams, n_samples;
    double resolution;
    const double *robot_pose;   // (W,3)
    const double *dirs;         // (n_beams,2) cos, sin computed on the host with the C library
    int map_kind, map_count, map_stride;
    const double *map_data;
    int32_t *hit;               // (W,n_beams)
    double2 *points;            // (W,n_beams)
};

// Exactly the reference's sample point and occupancy predicate (no FMA contraction):
// sensors/_lidar.py:113-115, world_map/_circles_world.py:47-50, _aabb_world.py:145-152.
__device__ __forceinline__ bool lidar_sample_occupied(const LidarArgs &a, const double *m, double cx, double cy,
                                                      double ca, double sa, int k, double &px, double &py)
{
    const double dist = __dmul_rn((double)k, a.resolution);
    px = __dadd_rn(cx, __dmul_rn(dist, ca));
    py = __dadd_rn(cy, __dmul_rn(dist, sa));
    if (a.map_kind == MAP_CIRCLES) {
        for (int c = 0; c < a.map_count; ++c) {
            const double dx = __dsub_rn(px, m[3 * c]), dy = __dsub_rn(py, m[3 * c + 1]);
            const double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            if (__dsub_rn(d, m[3 * c + 2]) <= 0.0) return true;
        }
    } else if (a.map_kind == MAP_AABB) {
        for (int c = 0; c < a.map_count; ++c) {
            const double tlx = m[4 * c], tly = m[4 * c + 1], bw = m[4 * c + 2], bh = m[4 * c + 3];
            if (px <= tlx && px >= __dsub_rn(tlx, bh) && py >= tly && py <= __dadd_rn(tly, bw)) return true;
        }
    }
    return false;
}

// One thread per (world, beam).  The reference marches n_samples points per beam through
// is_occupied; we obtain the same first-occupied sample index by bracketing it analytically
// (ray/circle, ray/box entry distance) and then testing the reference's exact sampled predicate in a
// window of a few samples around the entry -- O(C) per beam instead of O(C * n_samples).
__global__ void lidar_kernel(const __grid_constant__ LidarArgs a)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.W * a.n_beams) return;
    const int w = idx / a.n_beams, k = idx - w * a.n_beams;
    const double cx = a.robot_pose[3 * w], cy = a.robot_pose[3 * w + 1];
    const double ca = a.dirs[2 * k], sa = a.dirs[2 * k + 1];
    const double *m = a.map_data ? a.map_data + (size_t)w * a.map_stride : nullptr;
    int best = a.n_samples;
    double bx = 0.0, by = 0.0, px, py;
    if (a.map_kind != MAP_EMPTY && a.map_count > 0 && a.n_samples > 0) {
        if (lidar_sample_occupied(a, m, cx, cy, ca, sa, 0, px, py)) {
            best = -1;   // beam starts inside an obstacle: skipped (_lidar.py:118)
        } else {
            const double res = a.resolution;
            for (int c = 0; c < a.map_count && best != 1; ++c) {
                double t_in, t_out;
                if (a.map_kind == MAP_CIRCLES) {
                    const double ox = m[3 * c] - cx, oy = m[3 * c + 1] - cy, r = m[3 * c + 2];
                    const double tca = ox * ca + oy * sa;
                    const double d2 = fmax(ox * ox + oy * oy - tca * tca, 0.0);
                    const double rr = (r + 2.0 * res) * (r + 2.0 * res);   // generous: window, not decision
                    if (d2 > rr) continue;
                    const double thc = sqrt(rr - d2);
                    t_in = tca - thc; t_out = tca + thc;
                } else {
                    const double x_lo = m[4 * c] - m[4 * c + 3], x_hi = m[4 * c];
                    const double y_lo = m[4 * c + 1], y_hi = m[4 * c + 1] + m[4 * c + 2];
                    t_in = -CUDART_INF; t_out = CUDART_INF;
                    const double pad = 2.0 * res;
                    if (fabs(ca) > 1e-300) {
                        double t0 = (x_lo - pad - cx) / ca, t1 = (x_hi + pad - cx) / ca;
                        if (t0 > t1) { const double tmp = t0; t0 = t1; t1 = tmp; }
                        t_in = fmax(t_in, t0); t_out = fmin(t_out, t1);
                    } else if (cx < x_lo - pad || cx > x_hi + pad) continue;
                    if (fabs(sa) > 1e-300) {
                        double t0 = (y_lo - pad - cy) / sa, t1 = (y_hi + pad - cy) / sa;
                        if (t0 > t1) { const double tmp = t0; t0 = t1; t1 = tmp; }
                        t_in = fmax(t_in, t0); t_out = fmin(t_out, t1);
                    } else if (cy < y_lo - pad || cy > y_hi + pad) continue;
                    if (t_in > t_out) continue;
                }
                if (t_out < 0.0) continue;
                double k0d = floor(t_in / res) - 1.0, k1d = ceil(t_out / res) + 1.0;
                int k0 = k0d < 1.0 ? 1 : (k0d > 2.0e9 ? a.n_samples : (int)k0d);
                int k1 = k1d > (double)(best - 1) ? best - 1 : (int)k1d;
                for (int q = k0; q <= k1; ++q) {
                    if (lidar_sample_occupied(a, m, cx, cy, ca, sa, q, px, py)) {
                        best = q; bx = px; by = py;
                        break;
                    }
                }
            }
        }
    }
    const bool hit = best > 0 && best < a.n_samples;
    a.hit[idx] = hit ? best : -1;
    a.points[idx] = hit ? make_double2(bx, by) : make_double2(0.0, 0.0);
}

// OmniObstacleDetector: sensors/_omni_obstacle_detector.py:58-62
__global__ void omni_kernel(int W, const double *robot_pose, int map_kind, int map_count, int map_stride,
                            const double *map_data, double *out)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    const double *m = map_data ? map_data + (size_t)w * map_stride : nullptr;
    out[w] = map_closest(map_kind, map_count, m, robot_pose[3 * w], robot_pose[3 * w + 1]);
}

} // namespace pms
