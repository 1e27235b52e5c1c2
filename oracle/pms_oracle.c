/*
 * pms_oracle.c -- CPU restatement (fp64, plain C) of the pyminisim pedestrian-dynamics hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see pms_oracle.h).  Never shipped, never a fallback.
 *
 * Written from the behaviour of the reference (TimeEscaper/pyminisim); every function cites the
 * reference file:line it follows.  Floating-point operation order mirrors the reference where the
 * reference's own order is observable (so that the residual against the live reference is pure
 * libm / BLAS rounding, <= 1e-15 relative per step).
 */
#include "pms_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define ORC_PI 3.14159265358979323846
#define ORC_PED_RADIUS_CONST 0.3 /* core/_constants.py:2, used by the collision list */

/* ---- tiny pthread parallel-for over worlds (dynamic, one world per grab) ---- */
typedef void (*world_fn)(void *ctx, int w);
typedef struct { world_fn fn; void *ctx; int n; int next; } pf_job;
static void *pf_worker(void *arg)
{
    pf_job *j = (pf_job *)arg;
    for (;;) {
        int w = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (w >= j->n) break;
        j->fn(j->ctx, w);
    }
    return 0;
}
static void parallel_for_worlds(int n, int n_threads, world_fn fn, void *ctx)
{
    pf_job job = {fn, ctx, n, 0};
    if (n_threads > n) n_threads = n;
    if (n_threads <= 1) { pf_worker(&job); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)(n_threads - 1));
    int started = 0;
    for (int t = 0; t < n_threads - 1; ++t)
        if (pthread_create(&th[started], 0, pf_worker, &job) == 0) ++started;
    pf_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(th[t], 0);
    free(th);
}

/* Python / numpy float modulo: sign follows the divisor (numpy npy_divmod; numba lowers `%`
 * on float64 the same way).  Used at _hsfm.py:98,189 and util/_util.py:6. */
double orc_pymod(double a, double b)
{
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0.0) != (m < 0.0)) m += b;
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

/* util/_util.py:5-6, _hsfm.py:187-189 */
double orc_wrap_angle(double a) { return orc_pymod(a + ORC_PI, 2.0 * ORC_PI) - ORC_PI; }

/* pedestrians/_hsfm.py:113-134 (_calc_f_p_ij) */
void orc_pair_force(double radius_i, double radius_j, const double *r_i, const double *r_j,
                    const double *v_i, const double *v_j, double A, double B, double k1, double k2,
                    double *f_out)
{
    double r_ij = radius_i + radius_j;                 /* :124 */
    double dx = r_i[0] - r_j[0], dy = r_i[1] - r_j[1];
    double d_ij = sqrt(dx * dx + dy * dy);             /* :125 */
    double nx = dx / d_ij, ny = dy / d_ij;             /* :126 */
    double tx = -ny, ty = nx;                          /* :127 */
    double dv = (v_j[0] - v_i[0]) * tx + (v_j[1] - v_i[1]) * ty; /* :128 */
    double g = r_ij - d_ij;                            /* :129 */
    if (!(g > 0.0)) g = 0.0;
    double cn = A * exp((r_ij - d_ij) / B) + k1 * g;   /* :131 */
    double fx = cn * nx, fy = cn * ny;
    double ct = k2 * g * dv;                           /* :132 */
    fx += ct * tx;
    fy += ct * ty;
    f_out[0] = fx;
    f_out[1] = fy;
}

/* Forces and body-frame accelerations of pedestrians [64 blk, 64 blk + 64) of one world: _hsfm.py:146-158 and :46-110.
 * Pedestrians are independent within a step (each sums its own forces in j order), so blocks of rows can be shared among
 * threads without changing a bit of the result (large single crowds, orc_rollout with one world). */
typedef struct {
    const orc_batch *p; int n; const double *poses, *vels, *wps, *vmag, *robot4; double *dvb, *dom;
} row_ctx;
static void hsfm_rows(void *vctx, int blk)
{
    const row_ctx *q = (const row_ctx *)vctx;
    const orc_batch *p = q->p;
    const int n = q->n;
    const double *poses = q->poses, *vels = q->vels, *wps = q->wps, *vmag = q->vmag, *robot4 = q->robot4;
    double *dvb = q->dvb, *dom = q->dom;
    const double m = p->mass;
    const double I = 0.5 * (p->ped_radius * p->ped_radius); /* :239 */
    const int i_end = 64 * blk + 64 < n ? 64 * blk + 64 : n;
    for (int i = 64 * blk; i < i_end; ++i) {
        const double *ri = poses + 3 * i, *vi = vels + 3 * i;
        double theta = ri[2], omega = vi[2];
        /* _calc_desired_velocities :146-158 */
        double ddx = wps[2 * i] - ri[0], ddy = wps[2 * i + 1] - ri[1];
        double dn = sqrt(ddx * ddx + ddy * ddy);
        double vdx = ddx / dn * vmag[i], vdy = ddy / dn * vmag[i];
        /* f_0 :69-71 */
        double f0x = m * (vdx - vi[0]) / p->tau, f0y = m * (vdy - vi[1]) / p->tau;
        /* f_p :73-82, accumulation order j = 0..N-1 then robot */
        double fx = 0.0, fy = 0.0, f[2];
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            orc_pair_force(p->ped_radius, p->ped_radius, ri, poses + 3 * j, vi, vels + 3 * j,
                           p->A, p->B, p->k1, p->k2, f);
            fx += f[0];
            fy += f[1];
        }
        if (robot4) {
            orc_pair_force(p->ped_radius, p->hsfm_robot_radius, ri, robot4, vi, robot4 + 2,
                           p->A, p->B, p->k1, p->k2, f);
            fx += f[0];
            fy += f[1];
        }
        /* body frame :86-93 ; R = [[c,-s],[s,c]] (:269-270), r_f = R[:,0], r_o = R[:,1] */
        double c = cos(theta), s = sin(theta);
        double v_o = -s * vi[0] + c * vi[1];             /* (R^-1 v)[1] :88-89 */
        double u_f = (f0x + fx) * c + (f0y + fy) * s;    /* :91 */
        double u_o = p->ko * (fx * (-s) + fy * c) - p->kd * v_o; /* :92 */
        /* heading PD :95-100 */
        double klf = p->k_lambda * sqrt(f0x * f0x + f0y * f0y);
        double k_theta = I * klf;
        double k_omega = I * (1.0 + p->alpha) * sqrt(klf / p->alpha);
        double theta0 = orc_pymod(atan2(vdy, vdx), 2.0 * ORC_PI);
        double u_theta = -k_theta * orc_wrap_angle(theta - theta0) - k_omega * omega;
        dvb[2 * i] = 1.0 / m * u_f;                      /* :106 */
        dvb[2 * i + 1] = 1.0 / m * u_o;
        dom[i] = u_theta / I;                            /* :107 */
    }
}

static int g_row_threads = 1;   /* threads that share the rows of ONE world (set by orc_rollout when it has a single world) */

/* One world, one HSFM step without tracker: _hsfm.py:250-299 calling :146-158 and :46-110. */
void orc_hsfm_world_step(const orc_batch *p, int n, double *poses, double *vels, const double *wps,
                         const double *vmag, const double *robot4, double dt)
{
    double *dvb = (double *)malloc(sizeof(double) * 2 * (size_t)n);
    double *dom = (double *)malloc(sizeof(double) * (size_t)n);
    row_ctx rc = {p, n, poses, vels, wps, vmag, robot4, dvb, dom};
    parallel_for_worlds((n + 63) / 64, n >= 1024 ? g_row_threads : 1, hsfm_rows, &rc);
    /* explicit Euler :290-294 */
    for (int i = 0; i < n; ++i) {
        double *ri = poses + 3 * i, *vi = vels + 3 * i;
        double omega = vi[2];
        ri[0] = ri[0] + vi[0] * dt;
        ri[1] = ri[1] + vi[1] * dt;
        ri[2] = ri[2] + omega * dt;
        vi[2] = omega + dom[i] * dt;
        double c = cos(ri[2]), s = sin(ri[2]);
        double a = dvb[2 * i] * dt, b = dvb[2 * i + 1] * dt;
        vi[0] = vi[0] + (c * a + (-s) * b);
        vi[1] = vi[1] + (s * a + c * b);
    }
    free(dvb);
    free(dom);
}

/* ---------------- maps ---------------- */

static const double *map_of(const orc_batch *b, int w) { return b->map_data + (size_t)w * b->map_world_stride; }

/* world_map/_circles_world.py:47-50 (2-D branch), _aabb_world.py:145-152, _empty_world.py:27-30 */
static int map_is_occupied(const orc_batch *b, int w, double x, double y)
{
    const double *m = b->map_data ? map_of(b, w) : 0;
    if (b->map_kind == ORC_MAP_CIRCLES) {
        for (int c = 0; c < b->map_count; ++c) {
            double dx = x - m[3 * c], dy = y - m[3 * c + 1];
            if (sqrt(dx * dx + dy * dy) - m[3 * c + 2] <= 0.0) return 1;
        }
    } else if (b->map_kind == ORC_MAP_AABB) {
        for (int c = 0; c < b->map_count; ++c) {
            double tlx = m[4 * c], tly = m[4 * c + 1], bw = m[4 * c + 2], bh = m[4 * c + 3];
            if (x <= tlx && x >= (tlx - bh) && y >= tly && y <= (tly + bw)) return 1;
        }
    }
    return 0;
}

/* util/_geometry.py:4-42 restricted to one point, one segment */
static double point_segment_distance(double px, double py, double x1, double y1, double x2, double y2)
{
    double ex = x2 - x1, ey = y2 - y1;
    double ox = px - x1, oy = py - y1;
    double t = (ox * ex + oy * ey) / (ex * ex + ey * ey);
    if (t < 0.0) t = 0.0;
    if (t > 1.0) t = 1.0;
    double qx = x1 + t * ex, qy = y1 + t * ey;
    double dx = px - qx, dy = py - qy;
    return sqrt(dx * dx + dy * dy);
}

/* _circles_world.py:29-39, _aabb_world.py:98-113 (segments :31-35), _empty_world.py:19-25 */
static double map_closest_distance(const orc_batch *b, int w, double x, double y)
{
    double best = INFINITY;
    const double *m = b->map_data ? map_of(b, w) : 0;
    if (b->map_kind == ORC_MAP_CIRCLES) {
        for (int c = 0; c < b->map_count; ++c) {
            double dx = x - m[3 * c], dy = y - m[3 * c + 1];
            double d = sqrt(dx * dx + dy * dy) - m[3 * c + 2];
            if (d < best || d != d) best = d;
        }
    } else if (b->map_kind == ORC_MAP_AABB) {
        for (int c = 0; c < b->map_count; ++c) {
            double tlx = m[4 * c], tly = m[4 * c + 1], bw = m[4 * c + 2], bh = m[4 * c + 3];
            double seg[4][4] = {{tlx, tly, tlx, tly + bw},
                                {tlx - bh, tly, tlx - bh, tly + bw},
                                {tlx, tly, tlx - bh, tly},
                                {tlx, tly + bw, tlx - bh, tly + bw}};
            for (int k = 0; k < 4; ++k) {
                double d = point_segment_distance(x, y, seg[k][0], seg[k][1], seg[k][2], seg[k][3]);
                if (d < best || d != d) best = d;
            }
        }
    }
    return best;
}

/* core/_world_map.py:35-62 with eps_margin = 0.05; _empty_world.py:32-36 always False */
int orc_map_is_collision(const orc_batch *b, int w, double x, double y, double radius)
{
    if (b->map_kind == ORC_MAP_EMPTY || b->map_count == 0) return 0;
    int occ = map_is_occupied(b, w, x, y);
    double dist = map_closest_distance(b, w, x, y) - radius + 0.05;
    return occ || (dist <= 0.0);
}

/* ---------------- robot models ---------------- */

/* robot/_unicycle.py:53-69 (step) and :27-33 (velocity property) */
static void robot_step_unicycle(orc_batch *b, int w, const double *ctrl_in, double dt)
{
    double *pose = b->robot_pose + 3 * w, *ctrl = b->robot_ctrl + 2 * w, *vel = b->robot_vel + 3 * w;
    double c0 = ctrl_in ? ctrl_in[0] : ctrl[0], c1 = ctrl_in ? ctrl_in[1] : ctrl[1];
    double x = pose[0] + c0 * cos(pose[2]) * dt;
    double y = pose[1] + c0 * sin(pose[2]) * dt;
    double th = pose[2] + c1 * dt;
    th = orc_pymod(th + ORC_PI, 2.0 * ORC_PI) - ORC_PI;
    if (orc_map_is_collision(b, w, x, y, b->robot_radius)) {
        b->robot_world_collision[w] = 1; /* move rejected, state (incl. control) unchanged */
    } else {
        b->robot_world_collision[w] = 0;
        pose[0] = x; pose[1] = y; pose[2] = th;
        ctrl[0] = c0; ctrl[1] = c1;
    }
    vel[0] = ctrl[0] * cos(pose[2]);
    vel[1] = ctrl[0] * sin(pose[2]);
    vel[2] = ctrl[1];
}

/* robot/_bicycle.py:71-95 (step), :30-32 (velocity is always zero) */
static void robot_step_bicycle(orc_batch *b, int w, const double *ctrl_in, double dt)
{
    double *pose = b->robot_pose + 3 * w, *ctrl = b->robot_ctrl + 2 * w, *vel = b->robot_vel + 3 * w;
    double *rear = b->robot_rear + 3 * w;
    double v = ctrl_in ? ctrl_in[0] : ctrl[0], delta = ctrl_in ? ctrl_in[1] : ctrl[1];
    double xr = rear[0] + v * cos(rear[2]) * dt;
    double yr = rear[1] + v * sin(rear[2]) * dt;
    double th = rear[2] + (v / b->wheel_base) * tan(delta) * dt;
    th = orc_pymod(th + ORC_PI, 2.0 * ORC_PI) - ORC_PI;
    double d = b->wheel_base / 2.0;
    double xc = xr + d * cos(th), yc = yr + d * sin(th);
    if (orc_map_is_collision(b, w, xc, yc, b->robot_radius)) {
        b->robot_world_collision[w] = 1;
    } else {
        b->robot_world_collision[w] = 0;
        pose[0] = xc; pose[1] = yc; pose[2] = th;
        rear[0] = xr; rear[1] = yr; rear[2] = th;
        ctrl[0] = v; ctrl[1] = delta;
    }
    vel[0] = 0.0; vel[1] = 0.0; vel[2] = 0.0;
}

/* ---------------- sensors on the current state ---------------- */

/* core/_simulation.py:138-160 (collision list) and sensors/_pedestrian_detector.py:78-134 */
static void sense_world(orc_batch *b, int w, int accumulate)
{
    const int n = b->n_peds, words = (n + 31) / 32;
    uint32_t *cm = b->coll_mask ? b->coll_mask + (size_t)w * words : 0;
    uint32_t *dm = b->det_mask ? b->det_mask + (size_t)w * words : 0;
    if (cm) memset(cm, 0, sizeof(uint32_t) * words);
    if (dm) memset(dm, 0, sizeof(uint32_t) * words);
    if (b->robot_model == ORC_ROBOT_NONE) return;
    const double *rp = b->robot_pose + 3 * w;
    const double *poses = b->poses + (size_t)w * n * 3;
    for (int i = 0; i < n; ++i) {
        const double *pp = poses + 3 * i;
        double dx = rp[0] - pp[0], dy = rp[1] - pp[1];
        double distance = sqrt(dx * dx + dy * dy);
        if (cm && distance < b->robot_radius + ORC_PED_RADIUS_CONST) { /* _simulation.py:141-143 strict < */
            cm[i >> 5] |= 1u << (i & 31);
            if (accumulate && b->coll_count) b->coll_count[w]++;
        }
        if (!b->det_enabled || !dm) continue;
        double *out = b->det_vals + ((size_t)w * n + i) * 2;
        out[0] = 0.0; out[1] = 0.0;
        if (distance > b->det_max_dist) continue;                 /* :88 strict > */
        double point_angle = atan2(pp[1] - rp[1], pp[0] - rp[0]); /* :90 */
        double angle_diff = orc_wrap_angle(point_angle - rp[2]);  /* :91 */
        if (fabs(angle_diff) > b->det_fov / 2.0) continue;        /* :92 strict > */
        dm[i >> 5] |= 1u << (i & 31);
        if (accumulate && b->det_count) b->det_count[w]++;
        /* _convert_reading :114-134 */
        if (b->det_return_type == ORC_DET_POLAR) { out[0] = distance; out[1] = angle_diff; continue; }
        double xrr = distance * cos(angle_diff), yrr = distance * sin(angle_diff);
        if (b->det_return_type == ORC_DET_RELATIVE_ROTATED) { out[0] = xrr; out[1] = yrr; continue; }
        double th = rp[2];
        double xr = xrr * cos(th) - yrr * sin(th), yr = xrr * sin(th) + yrr * cos(th);
        if (b->det_return_type == ORC_DET_RELATIVE) { out[0] = xr; out[1] = yr; continue; }
        out[0] = xr + rp[0]; out[1] = yr + rp[1];
    }
}

static void sense_fn(void *ctx, int w) { sense_world((orc_batch *)ctx, w, 0); }

void orc_sense(orc_batch *b, int n_threads) { parallel_for_worlds(b->n_worlds, n_threads, sense_fn, b); }

/* ---------------- trackers ---------------- */

static int ev_cmp(const void *a, const void *b)
{
    const int *x = (const int *)a, *y = (const int *)b;
    for (int k = 0; k < 3; ++k) if (x[k] != y[k]) return x[k] < y[k] ? -1 : 1;
    return 0;
}

static void push_event(orc_batch *b, int step, int w, int i)
{
    if (!b->events || !b->ev_count) return;
    int slot = __atomic_fetch_add(b->ev_count, 1, __ATOMIC_RELAXED);
    if (slot < b->ev_cap) {
        b->events[3 * slot] = step; b->events[3 * slot + 1] = w; b->events[3 * slot + 2] = i;
    }
}

/* ---------------- fused rollout ---------------- */

static void rollout_world(orc_batch *b, int w, double dt, int n_steps)
{
    const int n = b->n_peds;
    double *poses = b->poses + (size_t)w * n * 3, *vels = b->vels + (size_t)w * n * 3;
    double *wps = b->cur_wp + (size_t)w * n * 2;
    const double *vmag = b->vmag + (size_t)w * n;
    double *backup = (double *)malloc(sizeof(double) * 3 * (size_t)n);
    for (int s = 0; s < n_steps; ++s) {
        /* core/_simulation.py:119-131: robot first, pedestrians see its post-step state */
        const double *ctrl_in = 0;
        if (b->controls && b->n_controls > 0) {
            int cs = s < b->n_controls ? s : b->n_controls - 1;
            ctrl_in = b->controls + ((size_t)cs * b->n_worlds + w) * 2;
        }
        if (b->robot_model == ORC_ROBOT_UNICYCLE) robot_step_unicycle(b, w, ctrl_in, dt);
        else if (b->robot_model == ORC_ROBOT_BICYCLE) robot_step_bicycle(b, w, ctrl_in, dt);
        double robot4[4];
        const double *r4 = 0;
        if (b->robot_model != ORC_ROBOT_NONE && b->robot_visible) { /* _hsfm.py:279-284 */
            robot4[0] = b->robot_pose[3 * w]; robot4[1] = b->robot_pose[3 * w + 1];
            robot4[2] = b->robot_vel[3 * w]; robot4[3] = b->robot_vel[3 * w + 1];
            r4 = robot4;
        }
        memcpy(backup, poses, sizeof(double) * 3 * (size_t)n); /* poses_backup, _hsfm.py:256 */
        orc_hsfm_world_step(b, n, poses, vels, wps, vmag, r4, dt);
        /* tracker on the NEW pose (_hsfm.py:301) */
        for (int i = 0; i < n; ++i) {
            size_t gi = (size_t)w * n + i;
            double dx = wps[2 * i] - poses[3 * i], dy = wps[2 * i + 1] - poses[3 * i + 1];
            if (!(sqrt(dx * dx + dy * dy) <= b->reach)) continue;
            if (b->tracker_kind == ORC_TRACKER_FIXED) {
                /* _fixed_waypoint_tracker.py:68-87 */
                int ni = b->wp_index[gi] + 1, steady = 0;
                if (ni >= b->n_table) {
                    if (b->loop) ni = 0; else { ni = b->n_table - 1; steady = 1; }
                }
                b->wp_index[gi] = ni;
                const double *t = b->table + (gi * b->n_table + ni) * 2;
                wps[2 * i] = t[0]; wps[2 * i + 1] = t[1];
                if (steady) { /* _hsfm.py:302-304: restore pre-step pose, zero velocity */
                    memcpy(poses + 3 * i, backup + 3 * i, sizeof(double) * 3);
                    vels[3 * i] = vels[3 * i + 1] = vels[3 * i + 2] = 0.0;
                }
            } else {
                /* _random_waypoint_tracker.py:66-91: current <- next[0], shift; the new tail is
                 * drawn by the host (global numpy RNG) when it replays the event log in order. */
                if (b->q_count[gi] > 0) {
                    const double *q = b->queue + (gi * b->q_cap + b->q_head[gi]) * 2;
                    wps[2 * i] = q[0]; wps[2 * i + 1] = q[1];
                    b->q_head[gi] = (b->q_head[gi] + 1) % b->q_cap;
                    b->q_count[gi]--;
                    push_event(b, s, w, i);
                } else {
                    push_event(b, -1 - s, w, i); /* starved: negative step marks the failure */
                }
            }
        }
        if (!b->nonfinite[w]) {
            int bad = 0;
            for (int k = 0; k < 3 * n; ++k) if (!isfinite(poses[k]) || !isfinite(vels[k])) bad = 1;
            if (bad) b->nonfinite[w] = (int)(b->steps_done + s + 1);
        }
        sense_world(b, w, 1);
    }
    free(backup);
}

typedef struct { orc_batch *b; double dt; int n_steps; } rollout_ctx;
static void rollout_fn(void *ctx, int w)
{
    rollout_ctx *c = (rollout_ctx *)ctx;
    rollout_world(c->b, w, c->dt, c->n_steps);
}

int orc_rollout(orc_batch *b, double dt, int n_steps, int n_threads)
{
    rollout_ctx c = {b, dt, n_steps};
    g_row_threads = b->n_worlds == 1 ? n_threads : 1;
    parallel_for_worlds(b->n_worlds, n_threads, rollout_fn, &c);
    g_row_threads = 1;
    if (b->events && b->ev_count) {
        int cnt = *b->ev_count < b->ev_cap ? *b->ev_count : b->ev_cap;
        qsort(b->events, (size_t)cnt, 3 * sizeof(int), ev_cmp);
    }
    b->steps_done += n_steps;
    return 0;
}

/* ---------------- lidar / omni ---------------- */

/* sensors/_lidar.py:105-129.  Brute force exactly as the reference: every sample of every beam is
 * tested with is_occupied; beam emits its first occupied sample if sample 0 is free. */
void orc_lidar_scan(const orc_batch *b, int n_beams, double max_dist, double resolution,
                    int32_t *hit_idx, double *points, int brute_force)
{
    (void)brute_force;
    int n_samples = (int)ceil((max_dist - 0.0) / resolution); /* np.arange length */
    for (int w = 0; w < b->n_worlds; ++w) {
        const double cx = b->robot_pose[3 * w], cy = b->robot_pose[3 * w + 1];
        for (int k = 0; k < n_beams; ++k) {
            /* np.linspace(0, 2*pi, n_beams): endpoint included and set exactly */
            double a;
            if (n_beams == 1) a = 0.0;
            else if (k == n_beams - 1) a = 2.0 * ORC_PI;
            else a = (double)k * ((2.0 * ORC_PI - 0.0) / (double)(n_beams - 1)) + 0.0;
            double ca = cos(a), sa = sin(a);
            int first = -1, occ0 = 0;
            double hx = 0.0, hy = 0.0;
            for (int s = 0; s < n_samples; ++s) {
                double dist = 0.0 + (double)s * resolution;
                double px = cx + dist * ca, py = cy + dist * sa;
                if (map_is_occupied(b, w, px, py)) {
                    if (s == 0) { occ0 = 1; break; }
                    first = s; hx = px; hy = py;
                    break;
                }
            }
            size_t o = (size_t)w * n_beams + k;
            if (occ0 || first < 0) { hit_idx[o] = -1; points[2 * o] = 0.0; points[2 * o + 1] = 0.0; }
            else { hit_idx[o] = first; points[2 * o] = hx; points[2 * o + 1] = hy; }
        }
    }
}

/* sensors/_omni_obstacle_detector.py:58-62 */
void orc_omni(const orc_batch *b, double *out)
{
    for (int w = 0; w < b->n_worlds; ++w)
        out[w] = map_closest_distance(b, w, b->robot_pose[3 * w], b->robot_pose[3 * w + 1]);
}
