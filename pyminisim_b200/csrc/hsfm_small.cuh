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
#include "pms_common.cuh"

namespace pms {

// robot state as seen by one step, double-buffered in shared memory
struct RobotSlot {
    double x, y, th;        // pose (fp64: detector / collision predicates)
    double vx, vy;          // linear velocity seen by the force term
};

struct RobotRegs {
    double x, y, th, c0, c1, rx, ry, rth, vx, vy, vw;
    int world_collision;
};

// Shared-memory image of one world.  The owner state lives here between steps (not in registers) so
// that nothing but the force accumulators is live across the pair loop.
//   spos {x,y}  svel {vx,vy}  st1 {theta,omega,wp_x,wp_y}  st2 {speed,cos,sin,wp_index}   state precision ST
//   ppos {x,y}  pvel {vx,vy}  in pair precision PT (alias spos/svel when PT == ST); plo = low parts of x,y (MIXED)
// Positions and velocities are separate arrays: the pair loop reads only positions (one LDS.64 per
// pair); velocities are touched only on body contact.
template <typename PT, typename ST, bool SPLIT>
struct SmallSmem {
    static constexpr bool ALIAS = sizeof(PT) == sizeof(ST);
    using ST2 = typename Vec2<ST>::type;
    using ST4 = typename Vec4<ST>::type;
    using PT2 = typename Vec2<PT>::type;
    __host__ __device__ static size_t off_svel(int Np) { return sizeof(ST2) * Np; }
    __host__ __device__ static size_t off_st1(int Np) { return 2 * sizeof(ST2) * Np; }
    __host__ __device__ static size_t off_st2(int Np) { return off_st1(Np) + sizeof(ST4) * Np; }
    __host__ __device__ static size_t end_state(int Np) { return off_st2(Np) + sizeof(ST4) * Np; }
    __host__ __device__ static size_t off_ppos(int Np) { return ALIAS ? 0 : end_state(Np); }
    __host__ __device__ static size_t off_pvel(int Np) { return ALIAS ? off_svel(Np) : end_state(Np) + sizeof(PT2) * Np; }
    __host__ __device__ static size_t off_plo(int Np) { return end_state(Np) + (ALIAS ? 0 : 2 * sizeof(PT2) * Np); }
    __host__ __device__ static size_t off_part(int Np) { return off_plo(Np) + (SPLIT ? sizeof(float2) * Np : 0); }
    __host__ __device__ static size_t off_robot(int Np, int S) { return off_part(Np) + (S > 1 ? sizeof(PT2) * Np * S : 0); }
    __host__ __device__ static size_t bytes(int Np, int S)
    {
        return (off_robot(Np, S) + 2 * sizeof(RobotSlot) + sizeof(RobotRegs) + 15) & ~size_t(15);
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
__device__ __noinline__ void robot_step(const HsfmArgs &a, int w, int s, RobotRegs &r)
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
        double sn, cs_;
        sincos(r.th, &sn, &cs_);
        const double nx = r.x + u0 * cs_ * a.dt;
        const double ny = r.y + u0 * sn * a.dt;
        double nth = r.th + u1 * a.dt;
        nth = py_mod(nth + PMS_PI, PMS_TWO_PI) - PMS_PI;
        if (map_is_collision(a.map_kind, a.map_count, map, nx, ny, a.robot_radius)) {
            r.world_collision = 1;            // move rejected: pose AND control stay (robot/_unicycle.py:64-65)
        } else {
            r.world_collision = 0;
            r.x = nx; r.y = ny; r.th = nth; r.c0 = u0; r.c1 = u1;
        }
        sincos(r.th, &sn, &cs_);
        r.vx = r.c0 * cs_; r.vy = r.c0 * sn; r.vw = r.c1;
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
        if (map_is_collision(a.map_kind, a.map_count, map, xc, yc, a.robot_radius)) {
            r.world_collision = 1;
        } else {
            r.world_collision = 0;
            r.x = xc; r.y = yc; r.th = nth;
            r.rx = xr; r.ry = yr; r.rth = nth;
            r.c0 = u0; r.c1 = u1;
        }
        r.vx = 0.0; r.vy = 0.0; r.vw = 0.0;   // BicycleRobotModelState.velocity is always zero
    }
}

// ---- per-pedestrian O(1) dynamics (pedestrians/_hsfm.py:69-71,86-110,146-158,290-294) ----------
// Advances one pedestrian by dt given the social force (fx,fy).  c,s = cos/sin(theta) cached from the
// previous step (R(theta') of step t is R(theta) of step t+1).
template <typename ST>
__device__ __forceinline__ void ped_integrate(const PedConsts<ST> &k, ST fx, ST fy, ST wx, ST wy, ST vm,
                                              ST &x, ST &y, ST &vx, ST &vy, ST &th, ST &om, ST &c, ST &s)
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
    x = x + vx * k.dt;
    y = y + vy * k.dt;
    th = th + om * k.dt;
    om = om + dom * k.dt;
    m_sincos(th, &s, &c);
    const ST a_ = dvbx * k.dt, b_ = dvby * k.dt;
    vx = vx + (c * a_ + (-s) * b_);
    vy = vy + (s * a_ + c * b_);
}

// ---- collision list + pedestrian detector for one pedestrian (fp64 predicates) -----------------
// core/_simulation.py:141-143 and sensors/_pedestrian_detector.py:86-103,114-134.
struct SenseOut {
    bool coll, det;
    double v0, v1;
};

// exact fp64 evaluation for a pedestrian that survived the cheap reject (rare -> out of line)
__device__ __noinline__ void sense_ped_exact(const HsfmArgs &a, const RobotSlot &rb, double px, double py,
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

// ---- the fused kernel -------------------------------------------------------------------------
// CTA-wide barrier 0.  The robot warp and the pedestrian warps run different loops and meet at the
// same hardware barrier from different instructions (warp-specialised producer / consumer).
__device__ __forceinline__ void cta_barrier() { asm volatile("bar.sync 0;" ::: "memory"); }

// One chunk of simulation steps [t_begin, t_end) for the wpc worlds of `group`.  Every thread of the
// CTA (pedestrian warps and the robot warp) executes the same sequence of CTA barriers.
template <typename PT, typename ST, bool SPLIT, int S>
__device__ __forceinline__ void small_chunk(const HsfmArgs &a, const int group, const int t_begin, const int t_end)
{
    using PT2 = typename Vec2<PT>::type;
    using ST2 = typename Vec2<ST>::type;
    using ST4 = typename Vec4<ST>::type;
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
    ST2 *spos = reinterpret_cast<ST2 *>(base);
    ST2 *svel = reinterpret_cast<ST2 *>(base + L::off_svel(Np));
    ST4 *st1 = reinterpret_cast<ST4 *>(base + L::off_st1(Np));
    ST4 *st2 = reinterpret_cast<ST4 *>(base + L::off_st2(Np));
    PT2 *ppos = reinterpret_cast<PT2 *>(base + L::off_ppos(Np));
    PT2 *pvel = reinterpret_cast<PT2 *>(base + L::off_pvel(Np));
    float2 *plo = reinterpret_cast<float2 *>(base + L::off_plo(Np));
    PT2 *part = reinterpret_cast<PT2 *>(base + L::off_part(Np));
    RobotSlot *rslot = reinterpret_cast<RobotSlot *>(base + L::off_robot(Np, S));
    RobotRegs &rr = *reinterpret_cast<RobotRegs *>(base + L::off_robot(Np, S) + 2 * sizeof(RobotSlot));

    const PairK<PT> &kpp = kpp_of(a, PT{});
    const PairK<PT> &kpr = kpr_of(a, PT{});
    const PedConsts<ST> &kc = kc_of(a, ST{});
    const size_t g = (size_t)w * N + i;

    // pair-phase view of the state (MIXED: hi/lo split of the fp64 position)
    auto publish = [&](ST x, ST y, ST vx, ST vy) {
        if constexpr (SPLIT) {
            const float xh = (float)x, yh = (float)y;
            ppos[i] = make_float2(xh, yh);
            pvel[i] = make_float2((float)vx, (float)vy);
            plo[i] = make_float2((float)(x - (double)xh), (float)(y - (double)yh));
        } else if constexpr (!L::ALIAS) {
            PT2 p; p.x = (PT)x; p.y = (PT)y; ppos[i] = p;
            PT2 v; v.x = (PT)vx; v.y = (PT)vy; pvel[i] = v;
        }
    };

    if (owner) {
        ST2 p0, v0; ST4 q1, q2;
        load_pv(a, g, p0.x, p0.y, v0.x, v0.y);
        const ST2 t = __ldcg(reinterpret_cast<const ST2 *>(a.th) + g);
        const ST2 q = __ldcg(reinterpret_cast<const ST2 *>(a.wp) + g);
        q1.x = t.x; q1.y = t.y; q1.z = q.x; q1.w = q.y;
        q2.x = reinterpret_cast<const ST *>(a.vmag)[g];
        m_sincos(t.x, &q2.z, &q2.y);
        q2.w = (ST)__ldcg(a.wp_index + g);        // small integer, exact in fp32
        spos[i] = p0; svel[i] = v0; st1[i] = q1; st2[i] = q2;
        publish(p0.x, p0.y, v0.x, v0.y);
    } else if (pair_warp && s == 0) {
        // padding lanes (N <= i < Np) run the pair loop for warp-uniform votes: park them far away
        PT2 far; far.x = (PT)1e18; far.y = (PT)(1e18 + 1e15 * i);
        PT2 z; z.x = 0; z.y = 0;
        ppos[i] = far; pvel[i] = z;
        if constexpr (SPLIT) plo[i] = make_float2(0.f, 0.f);
    }

    // ---- robot warp: its own loop (same barrier sequence), one step ahead of the pedestrians ----
    if (is_robot_warp) {
        const bool robot_lane = world_ok && has_robot;
        if (robot_lane) {
            robot_load(a, w, rr);
            robot_step(a, w, t_begin, rr);
            rslot[t_begin & 1] = RobotSlot{rr.x, rr.y, rr.th, rr.vx, rr.vy};
        }
        for (int t = t_begin; t < t_end; ++t) {
            cta_barrier();   // A
            if (robot_lane && t + 1 < t_end) {
                robot_step(a, w, t + 1, rr);
                rslot[(t + 1) & 1] = RobotSlot{rr.x, rr.y, rr.th, rr.vx, rr.vy};
            }
            cta_barrier();   // B
        }
        if (robot_lane) robot_store(a, w, rr);
        return;
    }

    unsigned det_acc = 0, coll_acc = 0;   // per owner warp
    int first_bad = 0;
    const int lane = tid & 31;

    for (int t = t_begin; t < t_end; ++t) {
        cta_barrier();     // A: state t and robot slot t&1 are visible
        const RobotSlot &rb = rslot[t & 1];
        PT fx = 0, fy = 0;
        if (pair_warp) {
            const PT2 me = ppos[i];
            float2 melo = make_float2(0.f, 0.f);
            if constexpr (SPLIT) melo = plo[i];
            PT fx1 = 0, fy1 = 0;
            int j = s;
            if constexpr (sizeof(PT) == 4) {
                // fp32 pair loop, 4 pairs per trip; one warp vote decides whether anybody is in contact
                for (; j + 3 * S < N; j += 4 * S) {
                    float dx[4], dy[4], inv[4], sg[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 o = ppos[j + q * S];
                        dx[q] = me.x - o.x; dy[q] = me.y - o.y;
                        if constexpr (SPLIT) { const float2 l = plo[j + q * S]; dx[q] += melo.x - l.x; dy[q] += melo.y - l.y; }
                    }
                    pair_far(kpp, dx[0], dy[0], j == i, inv[0], sg[0], fx, fy);
                    pair_far(kpp, dx[1], dy[1], j + S == i, inv[1], sg[1], fx1, fy1);
                    pair_far(kpp, dx[2], dy[2], j + 2 * S == i, inv[2], sg[2], fx, fy);
                    pair_far(kpp, dx[3], dy[3], j + 3 * S == i, inv[3], sg[3], fx1, fy1);
                    const float smax = fmaxf(fmaxf(sg[0], sg[1]), fmaxf(sg[2], sg[3]));
                    if (__any_sync(0xffffffffu, smax > 0.0f)) {
                        const float2 mv = pvel[i];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float2 ov = pvel[j + q * S];
                            pair_contact(kpp, dx[q], dy[q], ov.x - mv.x, ov.y - mv.y, inv[q], sg[q], fx, fy);
                        }
                    }
                }
                for (; j < N; j += S) {
                    const float2 o = ppos[j];
                    float dx0 = me.x - o.x, dy0 = me.y - o.y;
                    if constexpr (SPLIT) { const float2 l = plo[j]; dx0 += melo.x - l.x; dy0 += melo.y - l.y; }
                    const float2 mv = pvel[i], ov = pvel[j];
                    pair_accum(kpp, dx0, dy0, ov.x - mv.x, ov.y - mv.y, j == i, fx, fy);
                }
            } else {
                const PT2 mv = pvel[i];
                for (; j < N; j += S) {
                    const PT2 o = ppos[j], ov = pvel[j];
                    pair_accum(kpp, me.x - o.x, me.y - o.y, ov.x - mv.x, ov.y - mv.y, j == i, fx, fy);
                }
            }
            fx += fx1; fy += fy1;
            if (s == S - 1 && has_robot && a.robot_visible) {   // robot term (_hsfm.py:80-82)
                const PT2 mv = pvel[i];
                PT dx, dy;
                if constexpr (SPLIT) {
                    const float rxh = (float)rb.x, ryh = (float)rb.y;
                    dx = (me.x - rxh) + (melo.x - (float)(rb.x - (double)rxh));
                    dy = (me.y - ryh) + (melo.y - (float)(rb.y - (double)ryh));
                } else {
                    dx = me.x - (PT)rb.x; dy = me.y - (PT)rb.y;
                }
                pair_accum(kpr, dx, dy, (PT)rb.vx - mv.x, (PT)rb.vy - mv.y, false, fx, fy);
            }
            if (S > 1) { PT2 p; p.x = fx; p.y = fy; part[s * Np + i] = p; }
        }
        cta_barrier();     // B: all reads of state t done; slice partials visible
        if (owner_warp) {
            SenseOut so{false, false, 0.0, 0.0};
            const bool last = (t == a.n_steps - 1);   // last step of the launch
            if (owner) {
                if (S > 1) {
                    fx = 0; fy = 0;
#pragma unroll
                    for (int q = 0; q < S; ++q) { const PT2 p = part[q * Np + i]; fx += p.x; fy += p.y; }
                }
                ST2 p0 = spos[i], v0 = svel[i];
                ST4 q1 = st1[i], q2 = st2[i];
                const ST x0 = p0.x, y0 = p0.y, th0 = q1.x;
                ped_integrate<ST>(kc, (ST)fx, (ST)fy, q1.z, q1.w, q2.x, p0.x, p0.y, v0.x, v0.y, q1.x, q1.y, q2.y, q2.z);
                // ---- waypoint tracker on the NEW pose (_hsfm.py:301) ----
                const ST ex = q1.z - p0.x, ey = q1.w - p0.y;
                if (m_sqrt(ex * ex + ey * ey) <= kc.reach) {
                    if (a.tracker_kind == TRACKER_FIXED) {       // _fixed_waypoint_tracker.py:68-87
                        int ni = (int)q2.w + 1;
                        bool steady = false;
                        if (ni >= a.n_table) {
                            if (a.loop) ni = 0; else { ni = a.n_table - 1; steady = true; }
                        }
                        q2.w = (ST)ni;
                        const ST2 nw = reinterpret_cast<const ST2 *>(a.table)[g * a.n_table + ni];
                        q1.z = nw.x; q1.w = nw.y;
                        if (steady) {                            // _hsfm.py:302-304
                            p0.x = x0; p0.y = y0; q1.x = th0; v0.x = 0; v0.y = 0; q1.y = 0;
                            m_sincos(q1.x, &q2.z, &q2.y);
                        }
                    } else {                                     // _random_waypoint_tracker.py:66-91
                        const int cnt = a.q_count[g];
                        const int slot = atomicAdd(a.ev_count, 1);
                        if (cnt > 0) {
                            const int head = a.q_head[g];
                            const ST2 nw = reinterpret_cast<const ST2 *>(a.queue)[g * a.q_cap + head];
                            q1.z = nw.x; q1.w = nw.y;
                            a.q_head[g] = (head + 1) % a.q_cap;
                            a.q_count[g] = cnt - 1;
                            if (slot < a.ev_cap) { a.events[3 * slot] = t; a.events[3 * slot + 1] = w; a.events[3 * slot + 2] = i; }
                        } else if (slot < a.ev_cap) {            // starved queue: negative step marks it
                            a.events[3 * slot] = -1 - t; a.events[3 * slot + 1] = w; a.events[3 * slot + 2] = i;
                        }
                    }
                }
                spos[i] = p0; svel[i] = v0; st1[i] = q1; st2[i] = q2;
                publish(p0.x, p0.y, v0.x, v0.y);
                if (!first_bad && !m_finite(p0.x + p0.y + v0.x + v0.y + q1.x + q1.y)) first_bad = (int)(a.steps_done + t + 1);
                // ---- collision list + detector on the new state ----
                if (has_robot) so = sense_ped<ST>(a, rb, p0.x, p0.y, last);
                if (last && has_robot && a.det_enabled) a.det_vals[g] = make_double2(so.v0, so.v1);
            }
            if (has_robot) {   // owner warps are whole warps: ballots are safe
                const unsigned dm = __ballot_sync(0xffffffffu, so.det);
                const unsigned cm = __ballot_sync(0xffffffffu, so.coll);
                det_acc += __popc(dm);
                coll_acc += __popc(cm);
                if (last && lane == 0) {
                    const int words = (N + 31) >> 5;
                    if ((i >> 5) < words) {
                        a.det_mask[(size_t)w * words + (i >> 5)] = dm;
                        a.coll_mask[(size_t)w * words + (i >> 5)] = cm;
                    }
                }
            }
        }
    }

    // ---- write back (owners read their own shared-memory entries: no barrier needed) ----
    if (owner) {
        const ST2 p0 = spos[i], v0 = svel[i];
        const ST4 q1 = st1[i], q2 = st2[i];
        store_pv(a, g, p0.x, p0.y, v0.x, v0.y);
        ST2 t2; t2.x = q1.x; t2.y = q1.y;
        reinterpret_cast<ST2 *>(a.th)[g] = t2;
        ST2 w2; w2.x = q1.z; w2.y = q1.w;
        reinterpret_cast<ST2 *>(a.wp)[g] = w2;
        a.wp_index[g] = (int)q2.w;
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
template <typename PT, typename ST, bool SPLIT, int S>
__device__ __forceinline__ void hsfm_small_body(const HsfmArgs &a)
{
    if (a.n_chunks <= 1) {
        small_chunk<PT, ST, SPLIT, S>(a, blockIdx.x, 0, a.n_steps);
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
        small_chunk<PT, ST, SPLIT, S>(a, group, t_begin, t_end);
        __threadfence();
        cta_barrier();
        if (threadIdx.x == 0)
            asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(a.group_done + group), "r"(chunk + 1) : "memory");
    }
}

// Two register budgets: <=256 compute threads/CTA (several CTAs per SM) and big single-world CTAs.
template <typename PT, typename ST, bool SPLIT, int S>
__global__ void __launch_bounds__(288, 2) hsfm_small_kernel(const __grid_constant__ HsfmArgs a)
{
    hsfm_small_body<PT, ST, SPLIT, S>(a);
}
template <typename PT, typename ST, bool SPLIT, int S>
__global__ void __launch_bounds__(1024, 1) hsfm_small_kernel_big(const __grid_constant__ HsfmArgs a)
{
    hsfm_small_body<PT, ST, SPLIT, S>(a);
}

} // namespace pms
