// sensor_fires.cuh -- the robot-side sensors (lidar, omni obstacle detector) and the hold-time scheduler INSIDE a fused
// multi-step launch.
//
// Reference: Simulation._update_sensors_state (core/_simulation.py:162-185) keeps one hold_time per sensor: every step
// hold += dt; when hold >= period the sensor is read and hold = 0 (with dt = 0.01 a 10 Hz sensor fires every ELEVEN steps:
// ten additions of 0.01 give 0.09999999999999999 < 0.1).  LidarSensor.get_reading (sensors/_lidar.py:105-129) and
// OmniObstacleDetector.get_reading (sensors/_omni_obstacle_detector.py:58-62) see the robot AFTER the step and the map, never
// the pedestrians -- so the warp that integrates the robot a block of steps ahead (warp_robot_block in hsfm_warp.cuh, the
// service warp of hsfm_duo.cuh, the robot warp of hsfm_small.cuh, robot_window_kernel of hsfm_large.cuh) also takes the
// readings of the firing steps of its block, off the pedestrians' critical path, and writes them to the records of the launch.
#pragma once
#include "pms_common.cuh"

namespace pms {

// ---- the scheduler, replayed for one launch ------------------------------------------------------------------------------
// One lane per sensor kind.  hold[kind] persists across launches (device memory); fire[kind][t] = record slot or -1;
// count[kind] = readings of this launch.  Same fp64 additions and comparison as the reference => the same firing steps.
struct SchedArgs {
    int n_steps;
    double dt;
    int enabled[SENSOR_KINDS];
    double period[SENSOR_KINDS];
    double *hold;                 // (SENSOR_KINDS)
    int *fire[SENSOR_KINDS];      // (n_steps) each: record slot of the reading after launch step t, or -1
    int *steps[SENSOR_KINDS];     // (n_steps) each: launch step of record k (the inverse table, for the caller)
    int *count;                   // (SENSOR_KINDS)
};
static __global__ void sched_kernel(const SchedArgs s)
{
    const int kind = threadIdx.x;
    if (kind >= SENSOR_KINDS || !s.enabled[kind]) return;
    double hold = s.hold[kind];
    const double period = s.period[kind];
    int n = 0;
    for (int t = 0; t < s.n_steps; ++t) {
        hold = __dadd_rn(hold, s.dt);                  // core/_simulation.py:177
        int slot = -1;
        if (hold >= period) { s.steps[kind][n] = t; slot = n++; hold = 0.0; }   // :178-180
        s.fire[kind][t] = slot;
    }
    s.hold[kind] = hold;
    s.count[kind] = n;
}

// ---- one lidar beam ------------------------------------------------------------------------------------------------------
struct LidarGeom {
    int n_samples;
    double resolution;
    int map_kind, map_count;
    const double *map;            // this world's obstacles
};

// Exactly the reference's sample point and occupancy predicate (no FMA contraction):
// sensors/_lidar.py:113-115, world_map/_circles_world.py:47-50, _aabb_world.py:145-152.
__device__ __forceinline__ void lidar_sample_point(const LidarGeom &a, double cx, double cy, double ca, double sa, int k, double &px, double &py)
{
    const double dist = __dmul_rn((double)k, a.resolution);
    px = __dadd_rn(cx, __dmul_rn(dist, ca));
    py = __dadd_rn(cy, __dmul_rn(dist, sa));
}
// the predicate of ONE obstacle c on the sample point (the map's predicate is the OR over its obstacles)
__device__ __forceinline__ bool lidar_point_in_obstacle(const LidarGeom &a, int c, double px, double py)
{
    const double *m = a.map;
    if (a.map_kind == MAP_CIRCLES) {
        const double dx = __dsub_rn(px, m[3 * c]), dy = __dsub_rn(py, m[3 * c + 1]);
        const double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        return __dsub_rn(d, m[3 * c + 2]) <= 0.0;
    }
    const double tlx = m[4 * c], tly = m[4 * c + 1], bw = m[4 * c + 2], bh = m[4 * c + 3];
    return px <= tlx && px >= __dsub_rn(tlx, bh) && py >= tly && py <= __dadd_rn(tly, bw);
}
__device__ __forceinline__ bool lidar_sample_occupied(const LidarGeom &a, double cx, double cy, double ca, double sa, int k,
                                                      double &px, double &py)
{
    lidar_sample_point(a, cx, cy, ca, sa, k, px, py);
    if (a.map_kind != MAP_CIRCLES && a.map_kind != MAP_AABB) return false;
    for (int c = 0; c < a.map_count; ++c)
        if (lidar_point_in_obstacle(a, c, px, py)) return true;
    return false;
}

// The reference marches n_samples points per beam through is_occupied and takes the first occupied one.  The map's predicate
// is the OR over its obstacles, so that sample is the minimum over the obstacles of "first sample inside obstacle c": per
// obstacle the kernel brackets the entry analytically (ray / circle, ray / box; a generous window, not a decision) and
// evaluates the reference's exact sampled predicate OF THAT OBSTACLE on the few samples of the window -- O(C) per beam instead
// of O(C * n_samples), and O(1) predicates per window sample instead of O(C).  Returns the index of the first occupied sample
// (point in bx, by) or -1: no hit, or the beam starts inside an obstacle (skipped, _lidar.py:118).
static __device__ __noinline__ int lidar_beam(const LidarGeom &a, double cx, double cy, double ca, double sa, double &bx, double &by)
{
    int best = a.n_samples;
    double px, py;
    bx = 0.0; by = 0.0;
    if (a.map_kind != MAP_EMPTY && a.map_count > 0 && a.n_samples > 0) {
        const double *m = a.map;
        const double res = a.resolution;
        // The windows are bracketed in fp32 (single MUFU sqrt / rcp, conservative slack well above the fp32 error of map-sized
        // coordinates) -- only the sampled predicate itself is fp64.  Very fine resolutions keep the fp64 brackets.
        const bool f32_windows = res >= 1e-3;
        const float caf = (float)ca, saf = (float)sa, resf = (float)res;
        bool start_inside = false;                                      // sample 0 = the robot position itself (_lidar.py:118)
        if (f32_windows && a.map_kind == MAP_CIRCLES) {
            for (int c = 0; c < a.map_count; ++c) {
                const float ox = (float)(m[3 * c] - cx), oy = (float)(m[3 * c + 1] - cy), r = (float)m[3 * c + 2];
                const float o2 = fmaf(ox, ox, oy * oy), rs = r + 1e-3f + 1e-6f * o2;
                if (o2 <= rs * rs && lidar_point_in_obstacle(a, c, cx, cy)) start_inside = true;     // exact test only near / inside
            }
        } else {
            start_inside = lidar_sample_occupied(a, cx, cy, ca, sa, 0, px, py);
        }
        if (start_inside) {
            best = -1;
        } else if (f32_windows && a.map_kind == MAP_CIRCLES) {
            // fp32 error of the bracket for map-sized coordinates ~1e-5 m; the padded radius moves the window start at least
            // 1e-3 m in front of the true entry, floor() - 1 covers the rounding of the division: 2-3 exact samples per hit
            const float padf = 1e-3f, inv_res = 1.0f / resf;
            for (int c = 0; c < a.map_count && best != 1; ++c) {
                const float ox = (float)(m[3 * c] - cx), oy = (float)(m[3 * c + 1] - cy), rp = (float)m[3 * c + 2] + padf;
                const float o2 = fmaf(ox, ox, oy * oy);
                const float tca = fmaf(ox, caf, oy * saf);
                const float d2 = fmaxf(o2 - tca * tca, 0.0f);
                const float rr = fmaf(rp, rp, 1e-3f + 1e-6f * o2);      // generous: window, not decision
                if (d2 > rr) continue;
                const float thc = sqrtf(rr - d2);
                const float t_in = tca - thc, t_out = tca + thc;
                if (t_out < 0.0f) continue;
                const float k0f = floorf(t_in * inv_res) - 1.0f, k1f = ceilf(t_out * inv_res) + 1.0f;
                const int k0 = k0f < 1.0f ? 1 : (k0f > 2.0e9f ? a.n_samples : (int)k0f);
                const int k1 = k1f > (float)(best - 1) ? best - 1 : (int)k1f;
                for (int q = k0; q <= k1; ++q) {
                    lidar_sample_point(a, cx, cy, ca, sa, q, px, py);
                    if (lidar_point_in_obstacle(a, c, px, py)) {            // every sample inside obstacle c lies in its window
                        best = q; bx = px; by = py;
                        break;
                    }
                }
            }
        } else {
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
                    lidar_sample_point(a, cx, cy, ca, sa, q, px, py);
                    if (lidar_point_in_obstacle(a, c, px, py)) {            // every sample inside obstacle c lies in its window
                        best = q; bx = px; by = py;
                        break;
                    }
                }
            }
        }
    }
    const bool hit = best > 0 && best < a.n_samples;
    if (!hit) { bx = 0.0; by = 0.0; }
    return hit ? best : -1;
}

// LidarSensorNoise.noisify_reading (sensors/_lidar.py:79-87) of one beam: dropped with probability drop_prob (hit index -1),
// else mean + L z added.  k = world * n_beams + beam, step = absolute step count after which the reading is taken.
__device__ __forceinline__ void lidar_noise_apply(const LidarNoise &nz, size_t k, long long step, int &hit, double &px, double &py)
{
    if (hit < 0) return;
    const uint4 b = noise_block(nz.seed, NOISE_STREAM_LIDAR, k, (unsigned long long)step);
    if (noise_uniform(b.x) < nz.drop_prob) { hit = -1; px = 0.0; py = 0.0; return; }
    const double2 z = noise_normal2(b.y, b.z);
    px += nz.mean_x + nz.l00 * z.x;
    py += nz.mean_y + (nz.l10 * z.x + nz.l11 * z.y);
}
// OmniObstacleDetector._noisify_reading (sensors/_omni_obstacle_detector.py:64-69)
__device__ __forceinline__ double omni_noise_apply(const OmniNoise &nz, int w, long long step, double d)
{
    const uint4 b = noise_block(nz.seed, NOISE_STREAM_OMNI, (unsigned long long)w, (unsigned long long)step);
    if (noise_uniform(b.x) < nz.p_miss) return CUDART_INF;
    return d + (nz.mu + nz.sigma * noise_normal2(b.y, b.z).x);
}

// The lidar reading of world w taken after launch step t with the robot at (x, y), into record `slot`: one warp, beams strided
// over the lanes.  Rare (every 11th step with the reference's defaults) and out of line.
static __device__ __noinline__ void lidar_fire(const HsfmArgs &a, int w, int t, int slot, double x, double y, int lane)
{
    const SensorPlan &sp = a.sp;
    const double *map = a.map_data ? a.map_data + (size_t)w * a.map_stride : nullptr;
    const long long step = a.steps_done + t + 1;
    const LidarGeom lg{sp.lidar_samples, sp.lidar_res, a.map_kind, a.map_count, map};
    const size_t base = ((size_t)slot * a.W + w) * sp.lidar_beams;
    for (int k = lane; k < sp.lidar_beams; k += 32) {
        double bx, by;
        int hit = lidar_beam(lg, x, y, sp.lidar_dirs[2 * k], sp.lidar_dirs[2 * k + 1], bx, by);
        if (sp.lidar_noise.enabled) lidar_noise_apply(sp.lidar_noise, (size_t)w * sp.lidar_beams + k, step, hit, bx, by);
        sp.lidar_hit[base + k] = hit;
        sp.lidar_pts[base + k] = make_double2(bx, by);
    }
}

// Readings of the firing steps among launch steps s0 .. s0 + n - 1 (n <= 32) whose robot poses are ring[0 .. n-1] (any slot
// type with fp64 members x, y).  Called by a whole warp.  Omni: lane k takes the reading of step s0 + k (the obstacle loop runs
// once for the whole block); lidar: one warp per firing step, beams over the lanes.
template <typename Slot>
static __device__ __noinline__ void sensor_fires_block(const HsfmArgs &a, int w, int s0, int n, const Slot *ring, int lane)
{
    const SensorPlan &sp = a.sp;
    if (sp.fire[SENSOR_OMNI]) {
        const int slot = lane < n ? sp.fire[SENSOR_OMNI][s0 + lane] : -1;
        if (slot >= 0) {
            const double *map = a.map_data ? a.map_data + (size_t)w * a.map_stride : nullptr;
            double d = map_closest(a.map_kind, a.map_count, map, ring[lane].x, ring[lane].y);
            if (sp.omni_noise.enabled) d = omni_noise_apply(sp.omni_noise, w, a.steps_done + s0 + lane + 1, d);
            sp.omni_dist[(size_t)slot * a.W + w] = d;
        }
    }
    if (sp.fire[SENSOR_LIDAR]) {
        const int slot = lane < n ? sp.fire[SENSOR_LIDAR][s0 + lane] : -1;
        unsigned fires = __ballot_sync(0xffffffffu, slot >= 0);
        while (fires) {                                                 // warp-uniform
            const int k = __ffs(fires) - 1;
            fires &= fires - 1;
            lidar_fire(a, w, s0 + k, __shfl_sync(0xffffffffu, slot, k), ring[k].x, ring[k].y, lane);
        }
    }
}

} // namespace pms
