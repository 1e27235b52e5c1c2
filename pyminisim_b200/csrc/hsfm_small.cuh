// hsfm_small.cuh -- fused multi-step HSFM kernel for worlds that fit one CTA (N <= 1024).
//
// One launch advances every world by n_steps iterations of
//     robot step -> HSFM forces + explicit Euler -> waypoint hand-off / freeze -> collision list + detector
// (core/_simulation.py:119-160, pedestrians/_hsfm.py:250-306) with the whole world state resident in
// registers / shared memory; HBM is touched once at entry and once at exit.
//
// Thread layout of a CTA:  [ wpc worlds ] x [ S j-slices ] x [ Np pedestrians ]  + one robot warp.
//   * thread (wl, s, i) accumulates the pair forces on pedestrian i of local world wl over
//     j = s, s+S, s+2S, ... reading {x,y,vx,vy} of j from shared memory (warp-wide broadcast);
//   * slice partials meet in shared memory; the s == 0 threads ("owners") hold the pedestrian state
//     in registers and do the O(1) body-frame / heading / integration / tracker / sensor work, so warps
//     with s > 0 skip it entirely instead of idling lanes;
//   * the robot warp runs ONE STEP AHEAD (the robot does not depend on pedestrians), double-buffered in
//     shared memory, so its fp64 trigonometry never sits on the critical path of a step.
#pragma once
#include <type_traits>
#include "pms_common.cuh"
#include "sensor_fires.cuh"

namespace pms {

// robot state as seen by one step, double-buffered in shared memory
struct RobotSlot {
    double x, y, th;        // pose (fp64: detector / collision predicates)
    double vx, vy;          // linear velocity seen by the force term
};

struct RobotRegs {
    double x, y, th, c0, c1, rx, ry, rth, vx, vy, vw;
    double cs, sn;          // cos / sin of th (kept current by robot_load / robot_step)
    int world_collision;
};

// Shared-memory image of one world: structure-of-arrays, one scalar array per field (conflict-free
// scalar access by owners, 16-byte vector access along j in the pair loop).
//   owner state (ST):  x y vx vy theta omega wp_x wp_y speed cos sin  + int wp_index
//   pair view   (PT):  xs ys vxs vys   (alias the owner x y vx vy when PT == ST)
//   MIXED only       :  xlo ylo  = low parts of the fp64 position (x = (float)x + lo)
//   S > 1            :  slice partial forces fxp[S][Np], fyp[S][Np]
//   robot            :  ring of ROBOT_RING RobotSlots written ahead by the robot warp + its registers
// The owner state lives here between steps (not in registers) so that nothing but the force
// accumulators is live across the pair loop.
constexpr int ROBOT_RING = 32;    // slots; the robot warp runs up to ROBOT_BLOCK steps ahead
constexpr int ROBOT_BLOCK = 16;

template <typename PT, typename ST, bool SPLIT>
struct SmallSmem {
    static constexpr bool ALIAS = sizeof(PT) == sizeof(ST);
    static constexpr int N_OWNER = 11;
    __host__ __device__ static size_t off_widx(int Np) { return sizeof(ST) * N_OWNER * Np; }
    __host__ __device__ static size_t off_pair(int Np) { return ALIAS ? 0 : off_widx(Np) + sizeof(int) * Np; }
    __host__ __device__ static size_t end_pair(int Np) { return off_widx(Np) + sizeof(int) * Np + (ALIAS ? 0 : 4 * sizeof(PT) * Np); }
    __host__ __device__ static size_t off_lo(int Np) { return end_pair(Np); }
    __host__ __device__ static size_t off_part(int Np) { return off_lo(Np) + (SPLIT ? 2 * sizeof(float) * Np : 0); }
    __host__ __device__ static size_t off_robot(int Np, int S) { return (off_part(Np) + (S > 1 ? 2 * sizeof(PT) * Np * S : 0) + 15) & ~size_t(15); }
    __host__ __device__ static size_t bytes(int Np, int S)
    {
        return (off_robot(Np, S) + ROBOT_RING * sizeof(RobotSlot) + sizeof(RobotRegs) + 15) & ~size_t(15);
    }
};

// ---- robot models, executed by one lane per world (fp64) ---------------------------------------
__device__ __forceinline__ void robot_load(const HsfmArgs &a, int w, RobotRegs &r)
{
    r.x = a.robot_pose[3 * w]; r.y = a.robot_pose[3 * w + 1]; r.th = a.robot_pose[3 * w + 2];
    r.c0 = a.robot_ctrl[2 * w]; r.c1 = a.robot_ctrl[2 * w + 1];
    r.rx = a.robot_rear[3 * w]; r.ry = a.robot_rear[3 * w + 1]; r.rth = a.robot_rear[3 * w + 2];
    r.vx = a.robot_vel[3 * w]; r.vy = a.robot_vel[3 * w + 1]; r.vw = a.robot_vel[3 * w + 2];
    r.world_collision = a.robot_world_collision[w];
    sincos(r.th, &r.sn, &r.cs);
}

__device__ __forceinline__ void robot_store(const HsfmArgs &a, int w, const RobotRegs &r)
{
    a.robot_pose[3 * w] = r.x; a.robot_pose[3 * w + 1] = r.y; a.robot_pose[3 * w + 2] = r.th;
    a.robot_ctrl[2 * w] = r.c0; a.robot_ctrl[2 * w + 1] = r.c1;
    a.robot_rear[3 * w] = r.rx; a.robot_rear[3 * w + 1] = r.ry; a.robot_rear[3 * w + 2] = r.rth;
    a.robot_vel[3 * w] = r.vx; a.robot_vel[3 * w + 1] = r.vy; a.robot_vel[3 * w + 2] = r.vw;
    a.robot_world_collision[w] = r.world_collision;
}

// One robot step `s` of world w (robot/_unicycle.py:53-69,27-33; robot/_bicycle.py:71-95,30-32).
// WARP: executed by a whole warp with a warp-uniform `r` -- the scalar arithmetic is the same on every lane, the map veto
// spreads the obstacles over the lanes (map_is_collision_warp); else by one lane.
template <bool WARP>
__device__ __forceinline__ void robot_step_impl(const HsfmArgs &a, int w, int s, RobotRegs &r, int lane)
{
    if (a.robot_model != ROBOT_UNICYCLE && a.robot_model != ROBOT_BICYCLE) return;
    double u0 = r.c0, u1 = r.c1;
    if (a.n_controls > 0) {
        const int cs = s < a.n_controls ? s : a.n_controls - 1;
        u0 = a.controls[((size_t)cs * a.W + w) * 2];
        u1 = a.controls[((size_t)cs * a.W + w) * 2 + 1];
    }
    const double *map = a.map_data ? a.map_data + (size_t)w * a.map_stride : nullptr;
    if (a.robot_model == ROBOT_UNICYCLE) {
        // r.cs, r.sn are sincos(r.th) already (robot_load / the previous step): one fp64 sincos per step, not two
        const double nx = r.x + u0 * r.cs * a.dt;
        const double ny = r.y + u0 * r.sn * a.dt;
        double nth = r.th + u1 * a.dt;
        nth = py_mod(nth + PMS_PI, PMS_TWO_PI) - PMS_PI;
        const bool blocked = WARP ? map_is_collision_warp(a.map_kind, a.map_count, map, nx, ny, a.robot_radius, lane)
                                  : map_is_collision(a.map_kind, a.map_count, map, nx, ny, a.robot_radius);
        if (blocked) {
            r.world_collision = 1;            // move rejected: pose AND control stay (robot/_unicycle.py:64-65)
        } else {
            r.world_collision = 0;
            r.x = nx; r.y = ny; r.th = nth; r.c0 = u0; r.c1 = u1;
            double sn, cs_;
            sincos(nth, &sn, &cs_);
            r.cs = cs_; r.sn = sn;
        }
        r.vx = r.c0 * r.cs; r.vy = r.c0 * r.sn; r.vw = r.c1;
    } else {
        double sn, cs_;
        sincos(r.rth, &sn, &cs_);
        const double xr = r.rx + u0 * cs_ * a.dt;
        const double yr = r.ry + u0 * sn * a.dt;
        double nth = r.rth + (u0 / a.wheel_base) * tan(u1) * a.dt;
        nth = py_mod(nth + PMS_PI, PMS_TWO_PI) - PMS_PI;
        const double d = a.wheel_base / 2.0;
        sincos(nth, &sn, &cs_);
        const double xc = xr + d * cs_, yc = yr + d * sn;
        const bool blocked = WARP ? map_is_collision_warp(a.map_kind, a.map_count, map, xc, yc, a.robot_radius, lane)
                                  : map_is_collision(a.map_kind, a.map_count, map, xc, yc, a.robot_radius);
        if (blocked) {
            r.world_collision = 1;
        } else {
            r.world_collision = 0;
            r.x = xc; r.y = yc; r.th = nth; r.cs = cs_; r.sn = sn;
            r.rx = xr; r.ry = yr; r.rth = nth;
            r.c0 = u0; r.c1 = u1;
        }
        r.vx = 0.0; r.vy = 0.0; r.vw = 0.0;   // BicycleRobotModelState.velocity is always zero
    }
}
static __device__ __noinline__ void robot_step(const HsfmArgs &a, int w, int s, RobotRegs &r) { robot_step_impl<false>(a, w, s, r, 0); }
static __device__ __noinline__ void robot_step_warp(const HsfmArgs &a, int w, int s, RobotRegs &r, int lane) { robot_step_impl<true>(a, w, s, r, lane); }

// ---- per-pedestrian O(1) dynamics (pedestrians/_hsfm.py:69-71,86-110,146-158,290-294) ----------
// Advances one pedestrian by dt given the social force (fx,fy).  c,s = cos/sin(theta) cached from the
// previous step (R(theta') of step t is R(theta) of step t+1).
template <typename ST>
__device__ __forceinline__ void ped_integrate(const PedConsts<ST> &k, ST fx, ST fy, ST wx, ST wy, ST vm,
                                              ST &x, ST &y, ST &vx, ST &vy, ST &th, ST &om, ST &c, ST &s,
                                              ST noise_x = ST(0), ST noise_y = ST(0))
{
    const ST ddx = wx - x, ddy = wy - y;
    ST vdx, vdy, f0x, f0y, klf, komega, dom, dvbx, dvby, werr;
    if constexpr (sizeof(ST) == 8) {
        // fp64: reference operation order (divisions kept)
        const ST dn = m_sqrt(ddx * ddx + ddy * ddy);
        vdx = ddx / dn * vm; vdy = ddy / dn * vm;
        f0x = k.m * (vdx - vx) / k.tau; f0y = k.m * (vdy - vy) / k.tau;
    } else {
        // fp32: 1/|w - r| from MUFU.RSQ + one Newton step (no IEEE divide / sqrt sequences)
        const ST n2 = ddx * ddx + ddy * ddy;
        ST ri = fast_rsqrt(n2);
        ri = ri * fmaf(-0.5f * n2 * ri, ri, 1.5f);
        const ST sc = vm * ri;
        vdx = ddx * sc; vdy = ddy * sc;
        f0x = k.m_over_tau * (vdx - vx); f0y = k.m_over_tau * (vdy - vy);
    }
    const ST vo = -s * vx + c * vy;
    const ST uf = (f0x + fx) * c + (f0y + fy) * s;
    const ST uo = k.ko * (fx * (-s) + fy * c) - k.kd * vo;
    klf = k.kl * m_sqrt(f0x * f0x + f0y * f0y);
    if constexpr (sizeof(ST) == 8) {
        const ST th0 = py_mod(m_atan2(vdy, vdx), ST(PMS_TWO_PI));
        werr = wrap_angle(th - th0);
    } else {
        // wrap(theta - atan2(v_d)) == atan2(sin(theta - theta_0), cos(theta - theta_0)) built from the cached
        // (cos, sin) of theta: one atan2f, no fmod (differs from the reference only at exactly +-pi)
        werr = m_atan2(s * vdx - c * vdy, c * vdx + s * vdy);
    }
    if constexpr (sizeof(ST) == 8) {
        const ST ktheta = k.I * klf;
        komega = k.I * k.one_alpha * m_sqrt(klf / k.alpha);
        const ST uth = -ktheta * werr - komega * om;
        dom = uth / k.I;
        dvbx = ST(1) / k.m * uf; dvby = ST(1) / k.m * uo;
    } else {
        // I cancels: u_theta / I = -klf*werr - (1+alpha)*sqrt(klf/alpha)*om
        komega = k.one_alpha * m_sqrt(klf * k.inv_alpha);
        dom = -klf * werr - komega * om;
        dvbx = k.inv_m * uf; dvby = k.inv_m * uo;
    }
    dvbx += noise_x; dvby += noise_y;
    x = x + vx * k.dt;
    y = y + vy * k.dt;
    th = th + om * k.dt;
    om = om + dom * k.dt;
    m_sincos(th, &s, &c);
    const ST a_ = dvbx * k.dt, b_ = dvby * k.dt;
    vx = vx + (c * a_ + (-s) * b_);
    vy = vy + (s * a_ + c * b_);
}

// Compensated sum hi + lo <- hi + p (the carry of the old lo is folded into p by the caller): s = fl(hi + p),
// lo = p - (s - hi) (Fast2Sum: exact when |hi| >= |p|, which the Euler increments satisfy except around zero crossings,
// where the loss is <= ulp(p)).
__device__ __forceinline__ void comp_add(float &hi, float &lo, float p)
{
    const float s = hi + p;
    lo = p - (s - hi);
    hi = s;
}

// ped_integrate<float> on a state carried as unevaluated fp32 sums hi + lo (PMS_FP32, see hsfm_warp.cuh): same
// arithmetic, compensated Euler sums, sincos(theta_hi) rotated to first order by theta_lo.
struct StateLo { float x, y, vx, vy, th, om; };
__device__ __forceinline__ void ped_integrate_comp(const PedConsts<float> &k, float fx, float fy, float wx, float wy, float vm,
                                                   float &x, float &y, float &vx, float &vy, float &th, float &om,
                                                   float &c, float &s, StateLo &lo, float noise_x, float noise_y)
{
    const float ddx = (wx - x) - lo.x, ddy = (wy - y) - lo.y;
    const float n2 = ddx * ddx + ddy * ddy;
    float ri = fast_rsqrt(n2);
    ri = ri * fmaf(-0.5f * n2 * ri, ri, 1.5f);
    const float sc = vm * ri;
    const float vdx = ddx * sc, vdy = ddy * sc;
    const float f0x = k.m_over_tau * ((vdx - vx) - lo.vx), f0y = k.m_over_tau * ((vdy - vy) - lo.vy);
    const float vo = -s * vx + c * vy;
    const float uf = (f0x + fx) * c + (f0y + fy) * s;
    const float uo = k.ko * (fx * (-s) + fy * c) - k.kd * vo;
    const float klf = k.kl * m_sqrt(f0x * f0x + f0y * f0y);
    const float werr = m_atan2(s * vdx - c * vdy, c * vdx + s * vdy);
    const float komega = k.one_alpha * m_sqrt(klf * k.inv_alpha);
    const float dom = -klf * werr - komega * om;
    const float dvbx = k.inv_m * uf + noise_x, dvby = k.inv_m * uo + noise_y;
    comp_add(x, lo.x, fmaf(vx, k.dt, fmaf(lo.vx, k.dt, lo.x)));
    comp_add(y, lo.y, fmaf(vy, k.dt, fmaf(lo.vy, k.dt, lo.y)));
    comp_add(th, lo.th, fmaf(om, k.dt, fmaf(lo.om, k.dt, lo.th)));
    comp_add(om, lo.om, fmaf(dom, k.dt, lo.om));
    m_sincos(th, &s, &c);
    const float s0 = s;
    s = fmaf(c, lo.th, s); c = fmaf(-s0, lo.th, c);
    const float a_ = dvbx * k.dt, b_ = dvby * k.dt;
    comp_add(vx, lo.vx, (c * a_ + (-s) * b_) + lo.vx);
    comp_add(vy, lo.vy, (s * a_ + c * b_) + lo.vy);
}

// ---- collision list + pedestrian detector for one pedestrian (fp64 predicates) -----------------
// core/_simulation.py:141-143 and sensors/_pedestrian_detector.py:86-103,114-134.
struct SenseOut {
    bool coll, det;
    double v0, v1;
};

// exact fp64 evaluation for a pedestrian that survived the cheap reject (rare -> out of line)
static __device__ __noinline__ void sense_ped_exact(const HsfmArgs &a, const RobotSlot &rb, double px, double py,
                                             bool want_values, SenseOut &o)
{
    const double coll_r = a.robot_radius + PMS_PED_RADIUS_CONST;
    const double dx = rb.x - px, dy = rb.y - py;
    const double distance = sqrt(dx * dx + dy * dy);
    o.coll = distance < coll_r;
    if (!a.det_enabled || distance > a.det_max_dist) return;
    const double point_angle = atan2(py - rb.y, px - rb.x);
    const double diff = wrap_angle(point_angle - rb.th);
    if (fabs(diff) > a.det_half_fov) return;
    o.det = true;
    if (!want_values) return;
    if (a.det_type == DET_POLAR) { o.v0 = distance; o.v1 = diff; return; }
    double sd, cd;
    sincos(diff, &sd, &cd);
    const double xrr = distance * cd, yrr = distance * sd;
    if (a.det_type == DET_RELATIVE_ROTATED) { o.v0 = xrr; o.v1 = yrr; return; }
    double st, ct;
    sincos(rb.th, &st, &ct);
    const double xr = xrr * ct - yrr * st, yr = xrr * st + yrr * ct;
    if (a.det_type == DET_RELATIVE) { o.v0 = xr; o.v1 = yr; return; }
    o.v0 = xr + rb.x; o.v1 = yr + rb.y;
}

template <typename ST>
__device__ __forceinline__ SenseOut sense_ped(const HsfmArgs &a, const RobotSlot &rb, ST px, ST py, bool want_values)
{
    SenseOut o{false, false, 0.0, 0.0};
    // cheap reject in state precision with a generous margin; the exact test is fp64
    const ST qx = (ST)rb.x - px, qy = (ST)rb.y - py;
    if (qx * qx + qy * qy <= (ST)a.sense_lim2) sense_ped_exact(a, rb, (double)px, (double)py, want_values, o);
    return o;
}

// Owner-side view of one world's shared-memory state (scalar SoA arrays, see SmallSmem).
template <typename ST> struct OwnerView {
    ST *x, *y, *vx, *vy, *th, *om, *wx, *wy, *vm, *c, *s;
    int *widx;
    float *lo = nullptr;     // compensated FP32 state: low parts x y vx vy th om, Np floats each (else NULL)
    int Np_ = 0;
    __device__ __forceinline__ explicit OwnerView(unsigned char *base, int Np, size_t off_widx)
    {
        Np_ = Np;
        x = reinterpret_cast<ST *>(base); y = x + Np; vx = x + 2 * Np; vy = x + 3 * Np; th = x + 4 * Np; om = x + 5 * Np;
        wx = x + 6 * Np; wy = x + 7 * Np; vm = x + 8 * Np; c = x + 9 * Np; s = x + 10 * Np;
        widx = reinterpret_cast<int *>(base + off_widx);
    }
};

// One owner step for pedestrian i of world w: O(1) dynamics + explicit Euler (_hsfm.py:86-110,290-294), waypoint
// tracker on the new pose (:301-304), state write-back to shared memory, collision list + detector predicate.
template <typename ST, bool SENS = false, typename Publish>
__device__ __forceinline__ SenseOut owner_update(const HsfmArgs &a, const OwnerView<ST> &ow, int i, size_t g, int w, int t,
                                                 ST fx, ST fy, const RobotSlot &rb, bool has_robot, bool last,
                                                 int &first_bad, Publish &&publish, int det_slot = -1)
{
    using ST2 = typename Vec2<ST>::type;
    const PedConsts<ST> &kc = kc_of(a, ST{});
    ST *o_x = ow.x, *o_y = ow.y, *o_vx = ow.vx, *o_vy = ow.vy, *o_th = ow.th, *o_om = ow.om;
    ST *o_wx = ow.wx, *o_wy = ow.wy, *o_vm = ow.vm, *o_c = ow.c, *o_s = ow.s;
    int *o_widx = ow.widx;
    SenseOut so{false, false, 0.0, 0.0};
    ST x = o_x[i], y = o_y[i], vx = o_vx[i], vy = o_vy[i], th = o_th[i], om = o_om[i];
    ST wx = o_wx[i], wy = o_wy[i], c = o_c[i], sn = o_s[i];
    const ST x0 = x, y0 = y, th0 = th;
    ST nzx = 0, nzy = 0;
    if (a.accel_noise) { nzx = (ST)a.accel_noise[2 * g]; nzy = (ST)a.accel_noise[2 * g + 1]; }
    else if (a.accel_noise_std > 0.0) { const double2 z = accel_noise_draw(a, g, a.steps_done + t); nzx = (ST)z.x; nzy = (ST)z.y; }
    StateLo lo{0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, lo0 = lo;
    bool comp = false;
    if constexpr (sizeof(ST) == 4) comp = ow.lo != nullptr;
    if constexpr (sizeof(ST) == 4) {
        if (comp) {
            const float *l = ow.lo + i; const int Np = ow.Np_;
            lo = StateLo{l[0], l[Np], l[2 * Np], l[3 * Np], l[4 * Np], l[5 * Np]}; lo0 = lo;
            ped_integrate_comp(kc, fx, fy, wx, wy, o_vm[i], x, y, vx, vy, th, om, c, sn, lo, nzx, nzy);
        } else {
            ped_integrate<ST>(kc, (ST)fx, (ST)fy, wx, wy, o_vm[i], x, y, vx, vy, th, om, c, sn, nzx, nzy);
        }
    } else {
        ped_integrate<ST>(kc, (ST)fx, (ST)fy, wx, wy, o_vm[i], x, y, vx, vy, th, om, c, sn, nzx, nzy);
    }
    // ---- waypoint tracker on the NEW pose (_hsfm.py:301) ----
    const ST ex = (wx - x) - (ST)lo.x, ey = (wy - y) - (ST)lo.y;
    bool reached;
    if constexpr (sizeof(ST) == 8) reached = m_sqrt(ex * ex + ey * ey) <= kc.reach;
    else reached = ex * ex + ey * ey <= kc.reach * kc.reach;
    if (reached) {
        if (a.tracker_kind == TRACKER_FIXED) {       // _fixed_waypoint_tracker.py:68-87
            int ni = o_widx[i] + 1;
            bool steady = false;
            if (ni >= a.n_table) {
                if (a.loop) ni = 0; else { ni = a.n_table - 1; steady = true; }
            }
            o_widx[i] = ni;
            const ST2 nw = reinterpret_cast<const ST2 *>(a.table)[g * a.n_table + ni];
            wx = nw.x; wy = nw.y;
            if (steady) {                            // _hsfm.py:302-304
                x = x0; y = y0; th = th0; vx = 0; vy = 0; om = 0;
                m_sincos(th, &sn, &c);
                if constexpr (sizeof(ST) == 4) {
                    if (comp) {
                        lo = StateLo{lo0.x, lo0.y, 0.f, 0.f, lo0.th, 0.f};
                        const float s0 = sn;
                        sn = fmaf(c, lo.th, sn); c = fmaf(-s0, lo.th, c);
                    }
                }
            }
        } else {                                     // _random_waypoint_tracker.py:66-91
            const int cnt = a.q_count[g];
            const int slot = atomicAdd(a.ev_count, 1);
            if (cnt > 0) {
                const int head = a.q_head[g];
                const ST2 nw = reinterpret_cast<const ST2 *>(a.queue)[g * a.q_cap + head];
                wx = nw.x; wy = nw.y;
                a.q_head[g] = (head + 1) % a.q_cap;
                a.q_count[g] = cnt - 1;
                if (slot < a.ev_cap) { a.events[3 * slot] = t; a.events[3 * slot + 1] = w; a.events[3 * slot + 2] = i; }
            } else if (slot < a.ev_cap) {            // starved queue: negative step marks it
                a.events[3 * slot] = -1 - t; a.events[3 * slot + 1] = w; a.events[3 * slot + 2] = i;
            }
        }
        o_wx[i] = wx; o_wy[i] = wy;
    }
    o_x[i] = x; o_y[i] = y; o_vx[i] = vx; o_vy[i] = vy; o_th[i] = th; o_om[i] = om; o_c[i] = c; o_s[i] = sn;
    if constexpr (sizeof(ST) == 4) {
        if (comp) {
            float *l = ow.lo + i; const int Np = ow.Np_;
            l[0] = lo.x; l[Np] = lo.y; l[2 * Np] = lo.vx; l[3 * Np] = lo.vy; l[4 * Np] = lo.th; l[5 * Np] = lo.om;
        }
    }
    publish(x, y, vx, vy);
    if (!first_bad && !m_finite(x + y + vx + vy + th + om)) first_bad = (int)(a.steps_done + t + 1);
    // ---- collision list + detector on the new state ----
    // det_slot >= 0: the detector fires after this step (sensor plan) -- its reading goes to record det_slot of the launch
    const bool record_vals = SENS && det_slot >= 0 && a.sp.det_vals != nullptr;
    const bool want = last || record_vals;
    if (has_robot) {
        if constexpr (sizeof(ST) == 4) {
            if (comp) so = sense_ped<double>(a, rb, (double)x + (double)lo.x, (double)y + (double)lo.y, want);
            else so = sense_ped<ST>(a, rb, x, y, want);
        } else {
            so = sense_ped<ST>(a, rb, x, y, want);
        }
    }
    if (last && has_robot && a.det_enabled) a.det_vals[g] = make_double2(so.v0, so.v1);
    if constexpr (SENS) {
        if (record_vals && has_robot) a.sp.det_vals[(size_t)det_slot * a.W * a.N + g] = make_double2(so.v0, so.v1);
    }
    return so;
}

// ---- the fused kernel -------------------------------------------------------------------------
// CTA-wide barrier 0.  The robot warp and the pedestrian warps run different loops and meet at the
// same hardware barrier from different instructions (warp-specialised producer / consumer).
__device__ __forceinline__ void cta_barrier() { asm volatile("bar.sync 0;" ::: "memory"); }

// Barrier over the S*Np threads of one world (named barrier 1 + wl; a single warp just re-converges).
__device__ __forceinline__ void world_barrier(int wl, int count)
{
    if (count == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" :: "r"(wl + 1), "r"(count) : "memory");
}

// fp32 far-field pair force for TWO pairs at once on the Blackwell packed-FP32 pipe (FFMA2 / FMUL2 /
// FADD2): 8 packed FMA-pipe instructions + 4 MUFU for 2 pairs.  Same arithmetic as pair_far.
// 1 / d and the overlap r_ij - d of two pairs (the part of pair_far2 a contact pass re-derives, bit for bit)
__device__ __forceinline__ void pair_geom2(const PairK<float> &k, float2 dx, float2 dy, float2 &inv, float2 &sg)
{
    const float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
    inv.x = fast_rsqrt(d2.x); inv.y = fast_rsqrt(d2.y);
    float2 d = __fmul2_rn(d2, inv);
#ifdef PMS_REFINE_SQRT   // off: measured to make no difference to parity (profiles/r1_history.md), costs 9 %
    const float2 err = __ffma2_rn(make_float2(-d.x, -d.y), d, d2);
    d = __ffma2_rn(err, __fmul2_rn(inv, make_float2(0.5f, 0.5f)), d);
#endif
    sg = __fadd2_rn(make_float2(k.rij, k.rij), make_float2(-d.x, -d.y));
}
__device__ __forceinline__ void pair_far2(const PairK<float> &k, float2 dx, float2 dy, float2 &inv, float2 &sg,
                                          float2 &fx, float2 &fy)
{
    pair_geom2(k, dx, dy, inv, sg);
    const float2 arg = __fmul2_rn(sg, make_float2(k.cexp, k.cexp));
    float2 e; e.x = fast_ex2(arg.x); e.y = fast_ex2(arg.y);
    const float2 a = __fmul2_rn(e, __fmul2_rn(inv, make_float2(k.A, k.A)));
    fx = __ffma2_rn(a, dx, fx);
    fy = __ffma2_rn(a, dy, fy);
}

// One chunk of simulation steps [t_begin, t_end) for the wpc worlds of `group`.
// Synchronisation: the warps of one world meet twice per step at that world's named barrier; the robot
// warp fills the robot ring ROBOT_BLOCK steps ahead and everybody meets at CTA barrier 0 once per block.
// SENS: the instantiation that serves a sensor plan (pms_sensor_schedule); its code is compiled out of the default kernels
template <typename PT, typename ST, bool SPLIT, int S, bool SENS>
__device__ __forceinline__ void small_chunk(const HsfmArgs &a, const int group, const int t_begin, const int t_end)
{
    using ST2 = typename Vec2<ST>::type;
    using L = SmallSmem<PT, ST, SPLIT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int Np = a.Np, N = a.N;
    const int per_world = S * Np;
    const int n_compute = a.wpc * per_world;
    const int tid = threadIdx.x;
    const bool is_robot_warp = tid >= n_compute;
    const int wl = is_robot_warp ? (tid - n_compute) : tid / per_world;   // local world
    const int r = tid - wl * per_world;
    const int s = r / Np;
    const int i = r - s * Np;
    const int w = group * a.wpc + wl;
    const bool world_ok = (wl < a.wpc) && (w < a.W);
    const bool pair_warp = !is_robot_warp && world_ok;              // warp-uniform
    const bool active = pair_warp && i < N;
    const bool owner = active && s == 0;
    const bool owner_warp = pair_warp && s == 0;                     // warp-uniform
    const bool has_robot = a.robot_model != ROBOT_NONE;

    // carve the shared memory of local world wl
    unsigned char *base = smem_raw + (size_t)(wl < a.wpc ? wl : 0) * L::bytes(Np, S);
    OwnerView<ST> ow(base, Np, L::off_widx(Np));
    // compensated FP32 state: the low parts of all worlds of the CTA follow the per-world blocks (6 * Np floats each)
    if constexpr (sizeof(ST) == 4)
        if (a.comp) ow.lo = reinterpret_cast<float *>(smem_raw + (size_t)a.wpc * L::bytes(Np, S)) + (size_t)(wl < a.wpc ? wl : 0) * 6 * Np;
    ST *o_x = ow.x, *o_y = ow.y, *o_vx = ow.vx, *o_vy = ow.vy, *o_th = ow.th, *o_om = ow.om;
    ST *o_wx = ow.wx, *o_wy = ow.wy, *o_vm = ow.vm, *o_c = ow.c, *o_s = ow.s;
    int *o_widx = ow.widx;
    PT *xs = reinterpret_cast<PT *>(base + L::off_pair(Np));
    PT *ys = xs + Np, *vxs = xs + 2 * Np, *vys = xs + 3 * Np;
    float *xlo = reinterpret_cast<float *>(base + L::off_lo(Np));
    float *ylo = xlo + Np;
    PT *fxp = reinterpret_cast<PT *>(base + L::off_part(Np));
    PT *fyp = fxp + S * Np;
    RobotSlot *ring = reinterpret_cast<RobotSlot *>(base + L::off_robot(Np, S));
    RobotRegs &rr = *reinterpret_cast<RobotRegs *>(base + L::off_robot(Np, S) + ROBOT_RING * sizeof(RobotSlot));

    const PairK<PT> &kpp = kpp_of(a, PT{});
    const PairK<PT> &kpr = kpr_of(a, PT{});
    const PedConsts<ST> &kc = kc_of(a, ST{});
    const size_t g = (size_t)w * N + i;

    // pair-phase view of the state (MIXED: hi/lo split of the fp64 position)
    auto publish = [&](ST x, ST y, ST vx, ST vy) {
        if constexpr (SPLIT) {
            const float xh = (float)x, yh = (float)y;
            xs[i] = xh; ys[i] = yh; vxs[i] = (float)vx; vys[i] = (float)vy;
            xlo[i] = (float)(x - (double)xh); ylo[i] = (float)(y - (double)yh);
        } else if constexpr (!L::ALIAS) {
            xs[i] = (PT)x; ys[i] = (PT)y; vxs[i] = (PT)vx; vys[i] = (PT)vy;
        }
    };

    if (owner) {
        ST x, y, vx, vy, c, sn;
        load_pv(a, g, x, y, vx, vy);
        const ST2 t = __ldcg(reinterpret_cast<const ST2 *>(a.th) + g);
        const ST2 q = __ldcg(reinterpret_cast<const ST2 *>(a.wp) + g);
        m_sincos(t.x, &sn, &c);
        if constexpr (sizeof(ST) == 4) {
            if (ow.lo) {
                const float4 pl = __ldcg(reinterpret_cast<const float4 *>(a.pos_lo) + g);
                const float2 tl = __ldcg(reinterpret_cast<const float2 *>(a.th_lo) + g);
                float *l = ow.lo + i;
                l[0] = pl.x; l[Np] = pl.y; l[2 * Np] = pl.z; l[3 * Np] = pl.w; l[4 * Np] = tl.x; l[5 * Np] = tl.y;
                const float s0 = sn;
                sn = fmaf(c, tl.x, sn); c = fmaf(-s0, tl.x, c);
            }
        }
        o_x[i] = x; o_y[i] = y; o_vx[i] = vx; o_vy[i] = vy; o_th[i] = t.x; o_om[i] = t.y;
        o_wx[i] = q.x; o_wy[i] = q.y; o_vm[i] = reinterpret_cast<const ST *>(a.vmag)[g]; o_c[i] = c; o_s[i] = sn;
        o_widx[i] = __ldcg(a.wp_index + g);
        publish(x, y, vx, vy);
    } else if (owner_warp) {
        // padding lanes (N <= i < Np) run the pair loop so that warp votes stay uniform, and padded j
        // entries are read by the 4-wide pair loop: park them far away (their force underflows to 0)
        xs[i] = (PT)1e18; ys[i] = (PT)(1e18 + 1e15 * (i + 1)); vxs[i] = 0; vys[i] = 0;
        if constexpr (SPLIT) { xlo[i] = 0.f; ylo[i] = 0.f; }
    }

    // ---- robot warp: fills the ring one block ahead of the pedestrians ----
    if (is_robot_warp) {
        const bool robot_lane = world_ok && has_robot;
        int filled = t_begin;                  // ring holds steps < filled
        auto fill = [&](int upto) {
            const int from = filled;
            for (; filled < upto; ++filled) {       // `filled` advances on every lane: the sensor pass below is warp-uniform
                if (robot_lane) {
                    robot_step(a, w, filled, rr);
                    ring[(filled - t_begin) & (ROBOT_RING - 1)] = RobotSlot{rr.x, rr.y, rr.th, rr.vx, rr.vy};
                }
            }
            // lidar / omni readings of the firing steps just produced (sensor plan): lane wl stepped the robot of local world
            // wl, the whole warp reads the sensors of one world after the other
            if (SENS && a.sp.robot_on && has_robot) {
                __syncwarp();
                const int lane_ = tid & 31;
                for (int wq = 0; wq < a.wpc && group * a.wpc + wq < a.W; ++wq) {
                    const RobotSlot *rq = reinterpret_cast<const RobotSlot *>(smem_raw + (size_t)wq * L::bytes(Np, S) + L::off_robot(Np, S));
                    for (int st = from; st < upto; ++st)
                        sensor_fires_block(a, group * a.wpc + wq, st, 1, rq + ((st - t_begin) & (ROBOT_RING - 1)), lane_);
                }
                __syncwarp();
            }
        };
        if (robot_lane) robot_load(a, w, rr);
        fill(min(t_end, t_begin + ROBOT_BLOCK));
        for (int tb = t_begin; tb < t_end; tb += ROBOT_BLOCK) {
            cta_barrier();                     // block tb is complete; block tb - ROBOT_BLOCK is consumed
            fill(min(t_end, tb + 2 * ROBOT_BLOCK));
        }
        cta_barrier();                         // matches the pedestrians' final barrier
        if (robot_lane) robot_store(a, w, rr);
        return;
    }

    unsigned det_acc = 0, coll_acc = 0;   // per owner warp
    int first_bad = 0;
    const int lane = tid & 31;
    // contiguous j-slice of this thread: [j_lo, j_hi), multiples of 4
    const int N4 = (N + 3) & ~3;
    const int slice = (((N4 + S - 1) / S) + 3) & ~3;
    const int j_lo = min(N4, s * slice), j_hi = min(N4, j_lo + slice);
    const int i0 = i & ~31;               // first pedestrian of this warp: groups in [i0, i0+32) may hold j == i

    for (int t = t_begin; t < t_end; ++t) {
        if (((t - t_begin) & (ROBOT_BLOCK - 1)) == 0) cta_barrier();   // robot block hand-off
        if (pair_warp) world_barrier(wl, per_world);                    // A: state t visible to the world
        const RobotSlot &rb = ring[(t - t_begin) & (ROBOT_RING - 1)];
        PT fx = 0, fy = 0;
        if (pair_warp) {
            const PT mx = xs[i], my = ys[i];
            float mlx = 0.f, mly = 0.f;
            if constexpr (SPLIT) { mlx = xlo[i]; mly = ylo[i]; }
            if constexpr (sizeof(PT) == 4) {
                // ---- fp32 pair loop: 4 pairs per trip as two packed (f32x2) pairs ----
                float2 fxa = make_float2(0.f, 0.f), fya = fxa, fxb = fxa, fyb = fxa;
                const float2 m2x = make_float2(mx, mx), m2y = make_float2(my, my);
                const float2 l2x = make_float2(mlx, mlx), l2y = make_float2(mly, mly);
                // one group = 4 consecutive j: differences (MIXED: hi + lo), optional self mask
                auto load_group = [&](int jb, auto masked, float2 &dxa, float2 &dya, float2 &dxb, float2 &dyb) {
                    const float4 X = *reinterpret_cast<const float4 *>(xs + jb);
                    const float4 Y = *reinterpret_cast<const float4 *>(ys + jb);
                    dxa = __fadd2_rn(m2x, make_float2(-X.x, -X.y)); dxb = __fadd2_rn(m2x, make_float2(-X.z, -X.w));
                    dya = __fadd2_rn(m2y, make_float2(-Y.x, -Y.y)); dyb = __fadd2_rn(m2y, make_float2(-Y.z, -Y.w));
                    if constexpr (SPLIT) {
                        const float4 XL = *reinterpret_cast<const float4 *>(xlo + jb);
                        const float4 YL = *reinterpret_cast<const float4 *>(ylo + jb);
                        dxa = __fadd2_rn(dxa, __fadd2_rn(l2x, make_float2(-XL.x, -XL.y)));
                        dxb = __fadd2_rn(dxb, __fadd2_rn(l2x, make_float2(-XL.z, -XL.w)));
                        dya = __fadd2_rn(dya, __fadd2_rn(l2y, make_float2(-YL.x, -YL.y)));
                        dyb = __fadd2_rn(dyb, __fadd2_rn(l2y, make_float2(-YL.z, -YL.w)));
                    }
                    if constexpr (decltype(masked)::value) {
                        // j == i: move the partner far away => its term is exactly 0 (exp underflows)
                        const int k = i - jb;
                        dxa.x = k == 0 ? 1e18f : dxa.x; dxa.y = k == 1 ? 1e18f : dxa.y;
                        dxb.x = k == 2 ? 1e18f : dxb.x; dxb.y = k == 3 ? 1e18f : dxb.y;
                    }
                };
                auto contact_group = [&](int jb, float2 dxa, float2 dya, float2 dxb, float2 dyb, float2 inva, float2 invb,
                                         float2 sa, float2 sb) {
                    const float mvx = vxs[i], mvy = vys[i];
                    const float4 VX = *reinterpret_cast<const float4 *>(vxs + jb);
                    const float4 VY = *reinterpret_cast<const float4 *>(vys + jb);
                    pair_contact(kpp, dxa.x, dya.x, VX.x - mvx, VY.x - mvy, inva.x, sa.x, fx, fy);
                    pair_contact(kpp, dxa.y, dya.y, VX.y - mvx, VY.y - mvy, inva.y, sa.y, fx, fy);
                    pair_contact(kpp, dxb.x, dyb.x, VX.z - mvx, VY.z - mvy, invb.x, sb.x, fx, fy);
                    pair_contact(kpp, dxb.y, dyb.y, VX.w - mvx, VY.w - mvy, invb.y, sb.y, fx, fy);
                };
                auto groups = [&](int jb, const int je, auto masked) {
                    // Far-field pass over the whole range without a vote or branch between the trips (8 pairs per trip: four
                    // independent packed chains; consecutive trips overlap), then one vote on the largest overlap r_ij - d:
                    // the rare body-contact pass re-derives 1 / d and the overlap with the same operations, bit for bit.
                    const int j0 = jb;
                    float smax = -CUDART_INF_F;
                    for (; jb + 8 <= je; jb += 8) {
                        float2 dxa, dya, dxb, dyb, dxc, dyc, dxd, dyd, inva, invb, invc, invd, sa, sb, sc, sd;
                        load_group(jb, masked, dxa, dya, dxb, dyb);
                        load_group(jb + 4, masked, dxc, dyc, dxd, dyd);
                        pair_far2(kpp, dxa, dya, inva, sa, fxa, fya);
                        pair_far2(kpp, dxb, dyb, invb, sb, fxb, fyb);
                        pair_far2(kpp, dxc, dyc, invc, sc, fxa, fya);
                        pair_far2(kpp, dxd, dyd, invd, sd, fxb, fyb);
                        smax = fmaxf(smax, fmaxf(fmaxf(fmaxf(sa.x, sa.y), fmaxf(sb.x, sb.y)),
                                                 fmaxf(fmaxf(sc.x, sc.y), fmaxf(sd.x, sd.y))));
                    }
                    for (; jb < je; jb += 4) {
                        float2 dxa, dya, dxb, dyb, inva, invb, sa, sb;
                        load_group(jb, masked, dxa, dya, dxb, dyb);
                        pair_far2(kpp, dxa, dya, inva, sa, fxa, fya);
                        pair_far2(kpp, dxb, dyb, invb, sb, fxb, fyb);
                        smax = fmaxf(smax, fmaxf(fmaxf(sa.x, sa.y), fmaxf(sb.x, sb.y)));
                    }
                    if (!__any_sync(0xffffffffu, smax > 0.0f)) return;   // nobody in the warp touches anybody of this range
                    for (jb = j0; jb < je; jb += 4) {
                        float2 dxa, dya, dxb, dyb, inva, invb, sa, sb;
                        load_group(jb, masked, dxa, dya, dxb, dyb);
                        pair_geom2(kpp, dxa, dya, inva, sa);
                        pair_geom2(kpp, dxb, dyb, invb, sb);
                        const float s4 = fmaxf(fmaxf(sa.x, sa.y), fmaxf(sb.x, sb.y));
                        if (__any_sync(0xffffffffu, s4 > 0.0f)) contact_group(jb, dxa, dya, dxb, dyb, inva, invb, sa, sb);
                    }
                };
                const int m_lo = max(j_lo, min(j_hi, i0)), m_hi = min(j_hi, max(j_lo, i0 + 32));
                groups(j_lo, m_lo, std::false_type{});
                groups(m_lo, m_hi, std::true_type{});
                groups(m_hi, j_hi, std::false_type{});
                fx += (fxa.x + fxa.y) + (fxb.x + fxb.y);
                fy += (fya.x + fya.y) + (fyb.x + fyb.y);
            } else {
                const PT mvx = vxs[i], mvy = vys[i];
                for (int j = j_lo; j < min(j_hi, N); ++j)
                    pair_accum(kpp, mx - xs[j], my - ys[j], vxs[j] - mvx, vys[j] - mvy, j == i, fx, fy);
            }
            if (s == S - 1 && has_robot && a.robot_visible) {   // robot term (_hsfm.py:80-82)
                PT dx, dy;
                if constexpr (SPLIT) {
                    const float rxh = (float)rb.x, ryh = (float)rb.y;
                    dx = (mx - rxh) + (mlx - (float)(rb.x - (double)rxh));
                    dy = (my - ryh) + (mly - (float)(rb.y - (double)ryh));
                } else {
                    dx = mx - (PT)rb.x; dy = my - (PT)rb.y;
                }
                pair_accum(kpr, dx, dy, (PT)rb.vx - vxs[i], (PT)rb.vy - vys[i], false, fx, fy);
            }
            if (S > 1) { fxp[s * Np + i] = fx; fyp[s * Np + i] = fy; }
            world_barrier(wl, per_world);                               // B: reads of state t done; partials visible
        }
        if (owner_warp) {
            SenseOut so{false, false, 0.0, 0.0};
            const bool last = (t == a.n_steps - 1);   // last step of the launch
            int dslot = -1, cslot = -1;               // sensor plan: records of the detector / collision list after this step
            if constexpr (SENS) {
                if (a.sp.ped_on) {
                    if (a.sp.fire[SENSOR_DETECTOR]) dslot = a.sp.fire[SENSOR_DETECTOR][t];
                    if (a.sp.fire[SENSOR_COLLISIONS]) cslot = a.sp.fire[SENSOR_COLLISIONS][t];
                }
            }
            if (owner) {
                if (S > 1) {
                    fx = 0; fy = 0;
#pragma unroll
                    for (int q = 0; q < S; ++q) { fx += fxp[q * Np + i]; fy += fyp[q * Np + i]; }
                }
                so = owner_update<ST, SENS>(a, ow, i, g, w, t, (ST)fx, (ST)fy, rb, has_robot, last, first_bad, publish, dslot);
            }
            if (has_robot) {   // owner warps are whole warps: ballots are safe
                const unsigned dm = __ballot_sync(0xffffffffu, so.det);
                const unsigned cm = __ballot_sync(0xffffffffu, so.coll);
                det_acc += __popc(dm);
                coll_acc += __popc(cm);
                const int words = (N + 31) >> 5;
                if (lane == 0 && (i >> 5) < words) {
                    if (last) {
                        a.det_mask[(size_t)w * words + (i >> 5)] = dm;
                        a.coll_mask[(size_t)w * words + (i >> 5)] = cm;
                    }
                    if constexpr (SENS) {
                        if (dslot >= 0) a.sp.det_mask[((size_t)dslot * a.W + w) * words + (i >> 5)] = dm;
                        if (cslot >= 0) a.sp.coll_mask[((size_t)cslot * a.W + w) * words + (i >> 5)] = cm;
                    }
                }
            }
        }
    }
    cta_barrier();   // pairs with the robot warp's final barrier (robot ring may be re-used by the next chunk)

    // ---- write back (owners read their own shared-memory entries: no barrier needed) ----
    if (owner) {
        store_pv(a, g, o_x[i], o_y[i], o_vx[i], o_vy[i]);
        ST2 t2; t2.x = o_th[i]; t2.y = o_om[i];
        reinterpret_cast<ST2 *>(a.th)[g] = t2;
        ST2 w2; w2.x = o_wx[i]; w2.y = o_wy[i];
        reinterpret_cast<ST2 *>(a.wp)[g] = w2;
        a.wp_index[g] = o_widx[i];
        if constexpr (sizeof(ST) == 4) {
            if (ow.lo) {
                const float *l = ow.lo + i;
                reinterpret_cast<float4 *>(a.pos_lo)[g] = make_float4(l[0], l[Np], l[2 * Np], l[3 * Np]);
                reinterpret_cast<float2 *>(a.th_lo)[g] = make_float2(l[4 * Np], l[5 * Np]);
            }
        }
        if (first_bad) atomicMin(&a.nonfinite[w], first_bad);
    }
    if (owner_warp && has_robot && lane == 0) {
        if (det_acc) atomicAdd(&a.det_count[w], (unsigned long long)det_acc);
        if (coll_acc) atomicAdd(&a.coll_count[w], (unsigned long long)coll_acc);
    }
}

// Work distribution.  Static: CTA b runs all steps of group b.  Dynamic (n_chunks > 1): persistent CTAs
// pull (chunk, group) items in chunk-major order from a global counter; an item waits for the previous
// chunk of its group (already held by a running CTA, so the wait cannot deadlock).  With thousands of
// equal-length worlds this removes the partial last wave of a static grid (e.g. 1024 CTAs on 296 slots
// = 3.46 waves run as 4) at the price of one state round trip through HBM per chunk.
template <typename PT, typename ST, bool SPLIT, int S, bool SENS>
__device__ __forceinline__ void hsfm_small_body(const HsfmArgs &a)
{
    if (a.n_chunks <= 1) {
        small_chunk<PT, ST, SPLIT, S, SENS>(a, blockIdx.x, 0, a.n_steps);
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
        small_chunk<PT, ST, SPLIT, S, SENS>(a, group, t_begin, t_end);
        __threadfence();
        cta_barrier();
        if (threadIdx.x == 0)
            asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(a.group_done + group), "r"(chunk + 1) : "memory");
    }
}

// Two register budgets: <=256 compute threads/CTA (several CTAs per SM) and big single-world CTAs.
template <typename PT, typename ST, bool SPLIT, int S, bool SENS = false>
__global__ void __launch_bounds__(288, 2) hsfm_small_kernel(const __grid_constant__ HsfmArgs a)
{
    hsfm_small_body<PT, ST, SPLIT, S, SENS>(a);
}
template <typename PT, typename ST, bool SPLIT, int S, bool SENS = false>
__global__ void __launch_bounds__(1024, 1) hsfm_small_kernel_big(const __grid_constant__ HsfmArgs a)
{
    hsfm_small_body<PT, ST, SPLIT, S, SENS>(a);
}

} // namespace pms
