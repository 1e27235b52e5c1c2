// sensors.cuh -- stand-alone sensor kernels on the current state: collision list + pedestrian
// detector (no stepping), lidar ray march against the map, omni obstacle distance.
#pragma once
#include "hsfm_small.cuh"
#include "sensor_fires.cuh"

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
__global__ void lidar_noise_kernel(size_t nb, int32_t *hit, double2 *points, LidarNoise nz, long long step)
{
    const size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k >= nb) return;
    int h = hit[k];
    double2 p = points[k];
    lidar_noise_apply(nz, k, step, h, p.x, p.y);
    hit[k] = h; points[k] = p;
}

// OmniObstacleDetector._noisify_reading (sensors/_omni_obstacle_detector.py:64-69): inf with probability
// misdetection_prob, else distance + N(mu, sigma).
__global__ void omni_noise_kernel(int W, double *dist, OmniNoise nz, long long step)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    dist[w] = omni_noise_apply(nz, w, step, dist[w]);
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

// One thread per (world, beam): lidar_beam of sensor_fires.cuh on the stored robot pose (LidarSensor.get_reading,
// sensors/_lidar.py:105-129).
__global__ void lidar_kernel(const __grid_constant__ LidarArgs a)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.W * a.n_beams) return;
    const int w = idx / a.n_beams, k = idx - w * a.n_beams;
    const LidarGeom lg{a.n_samples, a.resolution, a.map_kind, a.map_count,
                       a.map_data ? a.map_data + (size_t)w * a.map_stride : nullptr};
    double bx, by;
    a.hit[idx] = lidar_beam(lg, a.robot_pose[3 * w], a.robot_pose[3 * w + 1], a.dirs[2 * k], a.dirs[2 * k + 1], bx, by);
    a.points[idx] = make_double2(bx, by);
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
