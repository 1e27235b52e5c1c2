// orca.cu -- ORCA (optimal reciprocal collision avoidance) agent step on the GPU.
//
// Replaces rvo2.PyRVOSimulator.doStep() + the adapter around it (pyminisim/pedestrians/_orca.py:95-127):
//   for every agent: nearest `maxNeighbors` agents within `neighborDist` (insertion-sorted by distance),
//   one ORCA half-plane per neighbour, 2-D incremental linear program (speed disc + half-planes) with
//   the "minimise the maximum penetration" fallback, then v <- v_new, p <- p + v dt; heading =
//   atan2(v); waypoint hand-off; preferred velocity towards the current waypoint (0 inside the
//   goal-reaching radius).
// Layout: one thread per agent, worlds packed several per CTA, positions/velocities of a world in
// shared memory (broadcast reads in the O(N) neighbour scan), n_steps fused in one launch with the
// agent state in registers.  fp32 like the library; this file is compiled with -fmad=false so every
// product and sum rounds separately, exactly like the library's scalar C++.
#include "orca.cuh"

#include <vector>

namespace pms {
namespace {

constexpr float kEps = 1e-5f;   // RVO_EPSILON

struct V2 { float x, y; };
__device__ __forceinline__ V2 mk(float x, float y) { V2 v; v.x = x; v.y = y; return v; }
__device__ __forceinline__ V2 operator+(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ V2 operator-(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ V2 operator-(V2 a) { return mk(-a.x, -a.y); }
__device__ __forceinline__ V2 operator*(float s, V2 a) { return mk(s * a.x, s * a.y); }
__device__ __forceinline__ V2 operator/(V2 a, float s) { const float inv = 1.0f / s; return mk(a.x * inv, a.y * inv); }
__device__ __forceinline__ float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float det(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ float norm2(V2 a) { return dot(a, a); }
__device__ __forceinline__ V2 unit(V2 a) { return a / sqrtf(norm2(a)); }

struct HalfPlane { V2 point, dir; };

// 1-D LP on the boundary of half-plane k, subject to the speed disc and half-planes 0..k-1
__device__ bool lp_line(const HalfPlane *hp, int k, float radius, V2 opt, bool direction_opt, V2 &result)
{
    const float dp = dot(hp[k].point, hp[k].dir);
    const float disc = dp * dp + radius * radius - norm2(hp[k].point);
    if (disc < 0.0f) return false;
    const float sq = sqrtf(disc);
    float t_lo = -dp - sq, t_hi = -dp + sq;
    for (int i = 0; i < k; ++i) {
        const float den = det(hp[k].dir, hp[i].dir);
        const float num = det(hp[i].dir, hp[k].point - hp[i].point);
        if (fabsf(den) <= kEps) {
            if (num < 0.0f) return false;
            continue;
        }
        const float t = num / den;
        if (den >= 0.0f) t_hi = fminf(t_hi, t);
        else t_lo = fmaxf(t_lo, t);
        if (t_lo > t_hi) return false;
    }
    if (direction_opt) {
        result = hp[k].point + (dot(opt, hp[k].dir) > 0.0f ? t_hi : t_lo) * hp[k].dir;
    } else {
        const float t = dot(hp[k].dir, opt - hp[k].point);
        const float tc = t < t_lo ? t_lo : (t > t_hi ? t_hi : t);
        result = hp[k].point + tc * hp[k].dir;
    }
    return true;
}

__device__ int lp_plane(const HalfPlane *hp, int n, float radius, V2 opt, bool direction_opt, V2 &result)
{
    if (direction_opt) result = radius * opt;
    else if (norm2(opt) > radius * radius) result = radius * unit(opt);
    else result = opt;
    for (int i = 0; i < n; ++i) {
        if (det(hp[i].dir, hp[i].point - result) > 0.0f) {
            const V2 keep = result;
            if (!lp_line(hp, i, radius, opt, direction_opt, result)) {
                result = keep;
                return i;
            }
        }
    }
    return n;
}

__device__ void lp_relax(const HalfPlane *hp, int n, int begin, float radius, V2 &result)
{
    float distance = 0.0f;
    HalfPlane proj[ORCA_MAX_NEIGHBORS];
    for (int i = begin; i < n; ++i) {
        if (det(hp[i].dir, hp[i].point - result) > distance) {
            int np = 0;
            for (int j = 0; j < i; ++j) {
                HalfPlane l;
                const float d = det(hp[i].dir, hp[j].dir);
                if (fabsf(d) <= kEps) {
                    if (dot(hp[i].dir, hp[j].dir) > 0.0f) continue;
                    l.point = 0.5f * (hp[i].point + hp[j].point);
                } else {
                    l.point = hp[i].point + (det(hp[j].dir, hp[i].point - hp[j].point) / d) * hp[i].dir;
                }
                l.dir = unit(hp[j].dir - hp[i].dir);
                proj[np++] = l;
            }
            const V2 keep = result;
            if (lp_plane(proj, np, radius, mk(-hp[i].dir.y, hp[i].dir.x), true, result) < np) result = keep;
            distance = det(hp[i].dir, hp[i].point - result);
        }
    }
}

struct OrcaArgs {
    int W, N, Np, wpc, n_steps, max_neighbors, n_table, loop;
    float time_step, neighbor_dist, time_horizon, radius;
    double goal_threshold, reach;
    float2 *pos, *vel, *pref;
    const float *max_speed;
    const double *pref_speed;
    double2 *cur_wp;
    const double2 *table;
    int *wp_index;
    double *heading;
    int *n_constraints;
};

__global__ void __launch_bounds__(256) orca_kernel(const __grid_constant__ OrcaArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Np = a.Np, N = a.N;
    const int wl = threadIdx.x / Np, i = threadIdx.x - wl * Np;
    const int w = blockIdx.x * a.wpc + wl;
    const bool valid = w < a.W && i < N;
    // per world: positions as two arrays sx | sy (four candidates per LDS.128, packed-FP32 differences), velocities float2 --
    // TWO copies by step parity: step s reads copy s & 1 and publishes the new state to the other one, so ONE barrier per
    // step is enough (a thread that runs ahead writes the copy nobody reads in this step).  Worlds are independent and never
    // share a barrier: a one-warp world (several per CTA) re-converges with __syncwarp, a larger world is its own CTA and uses
    // barrier 0.  (Round 2 measured why not named barriers 1 + local world with several worlds per CTA: a barrier id in a
    // register makes every CTA reserve all 16 hardware barriers, which caps the SM at 5 resident CTAs whatever their size --
    // C5 with 4 / 2 / 1 worlds per CTA took 2.59 / 3.68 / 6.41 ms that way and takes 2.40 with one world per CTA and barrier 0:
    // 14 instead of 16 worlds on the fullest SM.)
    float *sbase = reinterpret_cast<float *>(smem_raw) + (size_t)wl * 8 * Np;
    auto world_sync = [&]() {
        if (Np == 32) __syncwarp();
        else asm volatile("bar.sync 0;" ::: "memory");
    };
    const size_t g = (size_t)(valid ? w : 0) * N + (valid ? i : 0);

    V2 p = mk(0.f, 0.f), v = p, pref = p;
    float vmax = 0.f;
    double pspeed = 0.0, head = 0.0;
    double2 wp = make_double2(0.0, 0.0);
    int widx = 0, cnt = 0;
    if (valid) {
        const float2 pp = a.pos[g], vv = a.vel[g], pf = a.pref[g];
        p = mk(pp.x, pp.y); v = mk(vv.x, vv.y); pref = mk(pf.x, pf.y);
        vmax = a.max_speed[g]; pspeed = a.pref_speed[g]; wp = a.cur_wp[g]; widx = a.wp_index[g]; head = a.heading[g];
    }
    const int kmax = a.max_neighbors;
    const float inv_th = 1.0f / a.time_horizon;
    const float R = a.radius + a.radius, R2 = R * R;

    { sbase[i] = p.x; sbase[Np + i] = p.y; reinterpret_cast<float2 *>(sbase + 2 * Np)[i] = make_float2(v.x, v.y); }
    world_sync();
    for (int step = 0; step < a.n_steps; ++step) {
        const float *sx = sbase + (step & 1) * 4 * Np, *sy = sx + Np;
        const float2 *svel = reinterpret_cast<const float2 *>(sx + 2 * Np);
        V2 newv = v;
        if (valid) {
            // ---- neighbours: nearest kmax within neighbor_dist, insertion-sorted by squared distance ----
            float nd[ORCA_MAX_NEIGHBORS];
            int ni[ORCA_MAX_NEIGHBORS];
            cnt = 0;
            float range2 = a.neighbor_dist * a.neighbor_dist;
            if (kmax > 0) {
                // candidates in index order, four per trip (two LDS.128); the insertion logic of the library runs only for
                // the rare candidate inside the current range
                auto consider = [&](int o, float d2) {
                    if (o == i || o >= N || !(d2 < range2)) return;
                    if (cnt < kmax) ++cnt;
                    int k = cnt - 1;
                    while (k != 0 && d2 < nd[k - 1]) { nd[k] = nd[k - 1]; ni[k] = ni[k - 1]; --k; }
                    nd[k] = d2; ni[k] = o;
                    if (cnt == kmax) range2 = nd[cnt - 1];
                };
                const float2 px2 = make_float2(p.x, p.x), py2 = make_float2(p.y, p.y);
                for (int o = 0; o < N; o += 4) {
                    // |p - q|^2 = dx*dx + dy*dy with every operation rounded separately.  The final adds stay scalar: ptxas
                    // contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false, which would break the bit-exact
                    // match with the library's arithmetic.
                    const float4 X = *reinterpret_cast<const float4 *>(sx + o), Y = *reinterpret_cast<const float4 *>(sy + o);
                    const float2 dxa = __fadd2_rn(px2, make_float2(-X.x, -X.y)), dxb = __fadd2_rn(px2, make_float2(-X.z, -X.w));
                    const float2 dya = __fadd2_rn(py2, make_float2(-Y.x, -Y.y)), dyb = __fadd2_rn(py2, make_float2(-Y.z, -Y.w));
                    const float2 xa = __fmul2_rn(dxa, dxa), ya = __fmul2_rn(dya, dya), xb = __fmul2_rn(dxb, dxb), yb = __fmul2_rn(dyb, dyb);
                    const float2 da = make_float2(__fadd_rn(xa.x, ya.x), __fadd_rn(xa.y, ya.y));
                    const float2 db = make_float2(__fadd_rn(xb.x, yb.x), __fadd_rn(xb.y, yb.y));
                    if (fminf(fminf(da.x, da.y), fminf(db.x, db.y)) < range2) {
                        consider(o, da.x); consider(o + 1, da.y); consider(o + 2, db.x); consider(o + 3, db.y);
                    }
                }
            }
            // ---- one ORCA half-plane per neighbour ----
            HalfPlane hp[ORCA_MAX_NEIGHBORS];
            for (int k = 0; k < cnt; ++k) {
                const float2 qv = svel[ni[k]];
                const V2 rp = mk(sx[ni[k]], sy[ni[k]]) - p;
                const V2 rv = v - mk(qv.x, qv.y);
                const float d2 = norm2(rp);
                V2 u;
                HalfPlane l;
                if (d2 > R2) {
                    const V2 wv = rv - inv_th * rp;
                    const float w2 = norm2(wv);
                    const float dp1 = dot(wv, rp);
                    if (dp1 < 0.0f && dp1 * dp1 > R2 * w2) {
                        const float wl_ = sqrtf(w2);
                        const V2 uw = wv / wl_;
                        l.dir = mk(uw.y, -uw.x);
                        u = (R * inv_th - wl_) * uw;
                    } else {
                        const float leg = sqrtf(d2 - R2);
                        if (det(rp, wv) > 0.0f) l.dir = mk(rp.x * leg - rp.y * R, rp.x * R + rp.y * leg) / d2;
                        else l.dir = -(mk(rp.x * leg + rp.y * R, -rp.x * R + rp.y * leg) / d2);
                        const float dp2 = dot(rv, l.dir);
                        u = dp2 * l.dir - rv;
                    }
                } else {
                    const float inv_dt = 1.0f / a.time_step;
                    const V2 wv = rv - inv_dt * rp;
                    const float wl_ = sqrtf(norm2(wv));
                    const V2 uw = wv / wl_;
                    l.dir = mk(uw.y, -uw.x);
                    u = (R * inv_dt - wl_) * uw;
                }
                l.point = v + 0.5f * u;
                hp[k] = l;
            }
            const int failed = lp_plane(hp, cnt, vmax, pref, false, newv);
            if (failed < cnt) lp_relax(hp, cnt, failed, vmax, newv);
        }
        if (valid) {
            v = newv;
            p = mk(p.x + v.x * a.time_step, p.y + v.y * a.time_step);
            // the heading atan2(v) (_orca.py:103-104) is an output only: evaluated once after the last step
            // FixedWaypointTracker.update_waypoints on the new position (:106-107), steady flag ignored
            double dx = wp.x - (double)p.x, dy = wp.y - (double)p.y;
            double nr = sqrt(dx * dx + dy * dy);
            if (a.n_table > 0 && nr <= a.reach) {
                int nx = widx + 1;
                if (nx >= a.n_table) nx = a.loop ? 0 : a.n_table - 1;
                widx = nx;
                wp = a.table[g * a.n_table + nx];
                dx = wp.x - (double)p.x; dy = wp.y - (double)p.y;
                nr = sqrt(dx * dx + dy * dy);
            }
            // _update_preferred_velocities (:118-127)
            if (nr <= a.goal_threshold) pref = mk(0.f, 0.f);
            else pref = mk((float)(dx / nr * pspeed), (float)(dy / nr * pspeed));
        }
        {   // publish the new state in the other copy; one barrier per step
            float *nx_ = sbase + ((step + 1) & 1) * 4 * Np;
            nx_[i] = p.x; nx_[Np + i] = p.y; reinterpret_cast<float2 *>(nx_ + 2 * Np)[i] = make_float2(v.x, v.y);
        }
        world_sync();
    }
    if (valid) {
        if (a.n_steps > 0) head = atan2((double)v.y, (double)v.x);
        a.pos[g] = make_float2(p.x, p.y); a.vel[g] = make_float2(v.x, v.y); a.pref[g] = make_float2(pref.x, pref.y);
        a.cur_wp[g] = wp; a.wp_index[g] = widx; a.heading[g] = head; a.n_constraints[g] = cnt;
    }
}

// pref velocity from the current waypoint without stepping (ORCAPedestriansModel.__init__, _orca.py:86)
__global__ void orca_pref_kernel(const __grid_constant__ OrcaArgs a)
{
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= (size_t)a.W * a.N) return;
    const double2 wp = a.cur_wp[g];
    const float2 p = a.pos[g];
    const double dx = wp.x - (double)p.x, dy = wp.y - (double)p.y;
    const double nr = sqrt(dx * dx + dy * dy);
    if (nr <= a.goal_threshold) a.pref[g] = make_float2(0.f, 0.f);
    else a.pref[g] = make_float2((float)(dx / nr * a.pref_speed[g]), (float)(dy / nr * a.pref_speed[g]));
}

__global__ void orca_pack_kernel(size_t n, const double *pos, const double *vel, const double *ms, const double *ps,
                                 float2 *opos, float2 *ovel, float *oms, double *ops, double *heading)
{
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= n) return;
    opos[g] = make_float2((float)pos[2 * g], (float)pos[2 * g + 1]);
    ovel[g] = make_float2((float)vel[2 * g], (float)vel[2 * g + 1]);
    oms[g] = (float)ms[g];
    ops[g] = ps[g];
    heading[g] = 0.0;
}
__global__ void orca_unpack_kernel(size_t n, const float2 *pos, const float2 *vel, const double *heading, double *poses, double *vels)
{
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= n) return;
    poses[3 * g] = (double)pos[g].x; poses[3 * g + 1] = (double)pos[g].y; poses[3 * g + 2] = heading[g];
    vels[3 * g] = (double)vel[g].x; vels[3 * g + 1] = (double)vel[g].y; vels[3 * g + 2] = 0.0;   // _orca.py:101
}
__global__ void orca_gather_wp_kernel(size_t n, int n_table, const double2 *table, const int *index, double2 *wp)
{
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g < n) wp[g] = table[g * n_table + index[g]];
}
__global__ void orca_fill_kernel(size_t n, int *p, int v)
{
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g < n) p[g] = v;
}

inline unsigned nb(size_t n) { return (unsigned)((n + 255) / 256); }

OrcaArgs make_args(const OrcaState &s, int n_steps)
{
    OrcaArgs a{};
    a.W = s.W; a.N = s.N; a.Np = s.Np; a.n_steps = n_steps; a.max_neighbors = s.max_neighbors;
    a.n_table = s.n_table; a.loop = s.loop; a.time_step = s.time_step; a.neighbor_dist = s.neighbor_dist;
    a.time_horizon = s.time_horizon; a.radius = s.radius; a.goal_threshold = s.goal_threshold; a.reach = s.reach;
    a.pos = s.pos; a.vel = s.vel; a.pref = s.pref; a.max_speed = s.max_speed; a.pref_speed = s.pref_speed;
    a.cur_wp = s.cur_wp; a.table = s.table; a.wp_index = s.wp_index; a.heading = s.heading; a.n_constraints = s.n_constraints;
    a.wpc = s.Np == 32 ? 8 : 1;           // one-warp worlds: 8 per CTA; larger worlds: one CTA each (see world_sync in orca_kernel)
    if (a.wpc > s.W) a.wpc = s.W;
    return a;
}

int ensure_stage(OrcaState &s, size_t elems)
{
    if (elems <= s.stage_elems) return 0;
    if (s.stage) cudaFree(s.stage);
    s.stage = nullptr; s.stage_elems = 0;
    if (cudaMalloc((void **)&s.stage, elems * sizeof(double)) != cudaSuccess) return -2;
    s.stage_elems = elems;
    return 0;
}

} // namespace

int orca_alloc(OrcaState &s, int W, int N, double time_step, double neighbor_dist, int max_neighbors,
               double time_horizon, double radius, double goal_threshold)
{
    s.W = W; s.N = N; s.Np = (N + 31) & ~31;
    if (s.Np > 256) return -4;   // one thread per agent, one CTA per world group
    s.time_step = (float)time_step; s.neighbor_dist = (float)neighbor_dist; s.max_neighbors = max_neighbors;
    s.time_horizon = (float)time_horizon; s.radius = (float)radius; s.goal_threshold = goal_threshold;
    const size_t n = (size_t)W * N;
    bool ok = cudaMalloc((void **)&s.pos, n * sizeof(float2)) == cudaSuccess &&
              cudaMalloc((void **)&s.vel, n * sizeof(float2)) == cudaSuccess &&
              cudaMalloc((void **)&s.pref, n * sizeof(float2)) == cudaSuccess &&
              cudaMalloc((void **)&s.max_speed, n * sizeof(float)) == cudaSuccess &&
              cudaMalloc((void **)&s.pref_speed, n * sizeof(double)) == cudaSuccess &&
              cudaMalloc((void **)&s.cur_wp, n * sizeof(double2)) == cudaSuccess &&
              cudaMalloc((void **)&s.wp_index, n * sizeof(int)) == cudaSuccess &&
              cudaMalloc((void **)&s.heading, n * sizeof(double)) == cudaSuccess &&
              cudaMalloc((void **)&s.n_constraints, n * sizeof(int)) == cudaSuccess;
    if (!ok) return -2;
    cudaMemset(s.pref, 0, n * sizeof(float2));
    cudaMemset(s.n_constraints, 0, n * sizeof(int));
    cudaMemset(s.wp_index, 0, n * sizeof(int));
    return 0;
}

void orca_free(OrcaState &s)
{
    void *p[] = {s.pos, s.vel, s.pref, s.max_speed, s.pref_speed, s.cur_wp, s.table, s.wp_index, s.heading, s.n_constraints, s.stage};
    for (void *q : p) if (q) cudaFree(q);
    s = OrcaState{};
}

int orca_set_agents(OrcaState &s, const double *pos, const double *vel, const double *max_speeds,
                    const double *pref_speeds, cudaStream_t stream)
{
    const size_t n = (size_t)s.W * s.N;
    if (ensure_stage(s, n * 6)) return -2;
    double *d = s.stage;
    cudaMemcpyAsync(d, pos, n * 2 * sizeof(double), cudaMemcpyHostToDevice, stream);
    if (vel) cudaMemcpyAsync(d + 2 * n, vel, n * 2 * sizeof(double), cudaMemcpyHostToDevice, stream);
    else cudaMemsetAsync(d + 2 * n, 0, n * 2 * sizeof(double), stream);
    cudaMemcpyAsync(d + 4 * n, max_speeds, n * sizeof(double), cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(d + 5 * n, pref_speeds ? pref_speeds : max_speeds, n * sizeof(double), cudaMemcpyHostToDevice, stream);
    orca_pack_kernel<<<nb(n), 256, 0, stream>>>(n, d, d + 2 * n, d + 4 * n, d + 5 * n, s.pos, s.vel, s.max_speed, s.pref_speed, s.heading);
    s.agents_set = true;
    if (s.waypoints_set) orca_pref_kernel<<<nb(n), 256, 0, stream>>>(make_args(s, 0));
    return cudaStreamSynchronize(stream) == cudaSuccess ? 0 : -2;
}

int orca_set_waypoints(OrcaState &s, const double *table, int n_rows, const int32_t *start_index, double reach,
                       int loop, cudaStream_t stream)
{
    const size_t n = (size_t)s.W * s.N;
    if (s.table) { cudaFree(s.table); s.table = nullptr; }
    if (cudaMalloc((void **)&s.table, n * n_rows * sizeof(double2)) != cudaSuccess) return -2;
    cudaMemcpyAsync(s.table, table, n * n_rows * sizeof(double2), cudaMemcpyHostToDevice, stream);
    if (start_index) cudaMemcpyAsync(s.wp_index, start_index, n * sizeof(int), cudaMemcpyHostToDevice, stream);
    else orca_fill_kernel<<<nb(n), 256, 0, stream>>>(n, s.wp_index, n_rows > 1 ? 1 : 0);
    s.n_table = n_rows; s.reach = reach; s.loop = loop ? 1 : 0;
    orca_gather_wp_kernel<<<nb(n), 256, 0, stream>>>(n, n_rows, s.table, s.wp_index, s.cur_wp);
    s.waypoints_set = true;
    if (s.agents_set && s.refresh_pref) orca_pref_kernel<<<nb(n), 256, 0, stream>>>(make_args(s, 0));
    return cudaStreamSynchronize(stream) == cudaSuccess ? 0 : -2;
}

// Waypoints supplied by the caller (any tracker that runs on the host, e.g. RandomWaypointTracker): the device hand-off is
// switched off (n_table = 0) and the preferred velocities follow the given current waypoints (_orca.py:118-127).
int orca_set_current_waypoints(OrcaState &s, const double *current, cudaStream_t stream)
{
    const size_t n = (size_t)s.W * s.N;
    cudaMemcpyAsync(s.cur_wp, current, n * sizeof(double2), cudaMemcpyHostToDevice, stream);
    s.n_table = 0;
    s.waypoints_set = true;
    if (s.agents_set && s.refresh_pref) orca_pref_kernel<<<nb(n), 256, 0, stream>>>(make_args(s, 0));
    return cudaStreamSynchronize(stream) == cudaSuccess ? 0 : -2;
}

int orca_step(OrcaState &s, int n_steps, cudaStream_t stream, long long *launches)
{
    OrcaArgs a = make_args(s, n_steps);
    const int threads = a.wpc * s.Np;
    const int blocks = (s.W + a.wpc - 1) / a.wpc;
    const size_t smem = (size_t)a.wpc * 2 * 2 * s.Np * sizeof(float2);    // two copies (step parity) of sx | sy | svel per world
    orca_kernel<<<blocks, threads, smem, stream>>>(a);
    ++*launches;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int orca_get_agents(OrcaState &s, double *poses, double *vels, int32_t *n_constraints, cudaStream_t stream)
{
    const size_t n = (size_t)s.W * s.N;
    if (ensure_stage(s, n * 6)) return -2;
    orca_unpack_kernel<<<nb(n), 256, 0, stream>>>(n, s.pos, s.vel, s.heading, s.stage, s.stage + 3 * n);
    if (poses) cudaMemcpyAsync(poses, s.stage, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, stream);
    if (vels) cudaMemcpyAsync(vels, s.stage + 3 * n, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, stream);
    if (n_constraints) cudaMemcpyAsync(n_constraints, s.n_constraints, n * sizeof(int), cudaMemcpyDeviceToHost, stream);
    return cudaStreamSynchronize(stream) == cudaSuccess ? 0 : -2;
}

// poses | velocities (6 n doubles) into a pinned host block with ONE asynchronous copy; the caller synchronises
int orca_get_agents_async(OrcaState &s, double *pinned, cudaStream_t stream)
{
    const size_t n = (size_t)s.W * s.N;
    if (ensure_stage(s, n * 6)) return -2;
    orca_unpack_kernel<<<nb(n), 256, 0, stream>>>(n, s.pos, s.vel, s.heading, s.stage, s.stage + 3 * n);
    return cudaMemcpyAsync(pinned, s.stage, n * 6 * sizeof(double), cudaMemcpyDeviceToHost, stream) == cudaSuccess ? 0 : -2;
}

} // namespace pms
