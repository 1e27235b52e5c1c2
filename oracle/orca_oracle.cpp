/*
 * orca_oracle.cpp -- CPU restatement (fp32, C++) of the ORCA agent step used by
 * pyminisim/pedestrians/_orca.py:95-127.
 *
 * TEST INFRASTRUCTURE ONLY.  Never shipped, never a fallback.
 *
 * PARITY UNPINNED.  The arithmetic of this path lives in a third-party dependency that is NOT in
 * the reference tree: `rvo2` = Python-RVO2 (git+https://github.com/sybrenstuvel/Python-RVO2.git,
 * requirements.txt:7, no version or commit pin), a Cython wrapper around the RVO2 C++ library
 * v2.0.x (Agent / KdTree / RVOSimulator, single precision, RVO_EPSILON = 1e-5f).  It is not
 * installed in this image, cannot be downloaded (no network) and the reference holds no tests or
 * golden vectors at that boundary.  This file restates the *published* algorithm (van den Berg,
 * Guy, Lin, Manocha: "Reciprocal n-body collision avoidance", and the RVO2 library's documented
 * agent step: neighbour selection -> one ORCA half-plane per neighbour -> 2-D incremental LP with
 * the 3-D LP fallback -> v <- v_new, p <- p + v dt).  It is anchored on the reference's call
 * sites: PyRVOSimulator(timeStep, neighborDist, maxNeighbors, timeHorizon, timeHorizonObst,
 * radius, maxSpeed) _orca.py:73-79; addAgent(pos, maxSpeed=, velocity=) :82-84; doStep :99;
 * getAgentVelocity/getAgentPosition :101-104,120; setAgentPrefVelocity :124,127.
 *
 * Known, documented deviations from the library (all measure-zero on generic inputs):
 *   - neighbours are found by brute force in index order instead of a kd-tree; the selected set
 *     (the maxNeighbors nearest within neighborDist, sorted by distance) is identical except for
 *     exact distance ties, where the library's result depends on kd-tree traversal order.
 *   - no obstacles exist on this path (_orca.py never calls addObstacle), so obstacle half-planes
 *     and timeHorizonObst are omitted.
 *   - per-agent maxSpeed / initial velocity from addAgent(...) are honoured (the reference's
 *     evident intent); `honor_agent_kwargs = 0` reproduces a wrapper that ignores them.
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <atomic>
#include <thread>

namespace {

/* The algorithm is written once over the scalar type F.  F = float IS the oracle (rvo2 computes in single precision); the
 * F = double instantiation exists only for tests/test_oracle_orca_independent.py, which separates "is the algorithm right"
 * (fp64 instantiation against brute-force solvers, 1e-9) from "what does single precision cost" (fp32 against fp64). */
template <typename F> struct V2T {
    F x, y;
};
template <typename F> inline V2T<F> operator+(V2T<F> a, V2T<F> b) { return {a.x + b.x, a.y + b.y}; }
template <typename F> inline V2T<F> operator-(V2T<F> a, V2T<F> b) { return {a.x - b.x, a.y - b.y}; }
template <typename F> inline V2T<F> operator-(V2T<F> a) { return {-a.x, -a.y}; }
template <typename F> inline V2T<F> operator*(F s, V2T<F> a) { return {s * a.x, s * a.y}; }
template <typename F> inline V2T<F> operator/(V2T<F> a, F s) {
    const F inv = F(1) / s; /* the library divides through a reciprocal */
    return {a.x * inv, a.y * inv};
}
template <typename F> inline F dot(V2T<F> a, V2T<F> b) { return a.x * b.x + a.y * b.y; }
template <typename F> inline F det(V2T<F> a, V2T<F> b) { return a.x * b.y - a.y * b.x; }
template <typename F> inline F norm2(V2T<F> a) { return dot(a, a); }
template <typename F> inline F norm(V2T<F> a) { return std::sqrt(norm2(a)); }
template <typename F> inline V2T<F> unit(V2T<F> a) { return a / norm(a); }

template <typename F> struct HalfPlaneT {
    V2T<F> point, dir;
};

/* 1-D LP along half-plane `k`, constrained by the speed disc and half-planes 0..k-1. */
template <typename F>
bool lp_line(const HalfPlaneT<F> *hp, int k, F radius, V2T<F> opt, bool direction_opt, V2T<F> &result)
{
    const F kEps = F(1e-5f);
    const F dp = dot(hp[k].point, hp[k].dir);
    const F disc = dp * dp + radius * radius - norm2(hp[k].point);
    if (disc < F(0)) return false;
    const F sq = std::sqrt(disc);
    F t_lo = -dp - sq, t_hi = -dp + sq;
    for (int i = 0; i < k; ++i) {
        const F den = det(hp[k].dir, hp[i].dir);
        const F num = det(hp[i].dir, hp[k].point - hp[i].point);
        if (std::fabs(den) <= kEps) {
            if (num < F(0)) return false;
            continue;
        }
        const F t = num / den;
        if (den >= F(0)) t_hi = std::fmin(t_hi, t);
        else t_lo = std::fmax(t_lo, t);
        if (t_lo > t_hi) return false;
    }
    if (direction_opt) {
        result = hp[k].point + (dot(opt, hp[k].dir) > F(0) ? t_hi : t_lo) * hp[k].dir;
    } else {
        const F t = dot(hp[k].dir, opt - hp[k].point);
        const F tc = t < t_lo ? t_lo : (t > t_hi ? t_hi : t);
        result = hp[k].point + tc * hp[k].dir;
    }
    return true;
}

/* 2-D incremental LP. Returns the index of the first half-plane that made it infeasible, or n. */
template <typename F>
int lp_plane(const HalfPlaneT<F> *hp, int n, F radius, V2T<F> opt, bool direction_opt, V2T<F> &result)
{
    if (direction_opt) result = radius * opt;
    else if (norm2(opt) > radius * radius) result = radius * unit(opt);
    else result = opt;
    for (int i = 0; i < n; ++i) {
        if (det(hp[i].dir, hp[i].point - result) > F(0)) {
            const V2T<F> keep = result;
            if (!lp_line(hp, i, radius, opt, direction_opt, result)) {
                result = keep;
                return i;
            }
        }
    }
    return n;
}

/* Fallback: minimise the maximum penetration (no obstacle half-planes on this path). */
template <typename F>
void lp_relax(const HalfPlaneT<F> *hp, int n, int begin, F radius, V2T<F> &result)
{
    const F kEps = F(1e-5f);
    F distance = F(0);
    std::vector<HalfPlaneT<F>> proj;
    for (int i = begin; i < n; ++i) {
        if (det(hp[i].dir, hp[i].point - result) > distance) {
            proj.clear();
            for (int j = 0; j < i; ++j) {
                HalfPlaneT<F> l;
                const F d = det(hp[i].dir, hp[j].dir);
                if (std::fabs(d) <= kEps) {
                    if (dot(hp[i].dir, hp[j].dir) > F(0)) continue;
                    l.point = F(0.5) * (hp[i].point + hp[j].point);
                } else {
                    l.point = hp[i].point + (det(hp[j].dir, hp[i].point - hp[j].point) / d) * hp[i].dir;
                }
                l.dir = unit(hp[j].dir - hp[i].dir);
                proj.push_back(l);
            }
            const V2T<F> keep = result;
            if (lp_plane(proj.data(), (int)proj.size(), radius, V2T<F>{-hp[i].dir.y, hp[i].dir.x}, true, result) <
                (int)proj.size())
                result = keep;
            distance = det(hp[i].dir, hp[i].point - result);
        }
    }
}

/* The ORCA half-plane agent `a` takes on for neighbour `o`: u is the smallest change of the relative velocity that leaves
 * the velocity obstacle truncated at the time horizon (or, when the discs already overlap, at one time step); the agent
 * takes half of it (reciprocity): point = v_a + u / 2, the line runs along `dir` with the permitted side on its left. */
template <typename F>
HalfPlaneT<F> orca_halfplane(V2T<F> pa, V2T<F> po, V2T<F> va, V2T<F> vo, F radius, F inv_th, F time_step)
{
    const V2T<F> rp = po - pa;
    const V2T<F> rv = va - vo;
    const F d2 = norm2(rp);
    const F R = radius + radius;
    const F R2 = R * R;
    V2T<F> u;
    HalfPlaneT<F> l;
    if (d2 > R2) {
        const V2T<F> wv = rv - inv_th * rp;
        const F w2 = norm2(wv);
        const F dp1 = dot(wv, rp);
        if (dp1 < F(0) && dp1 * dp1 > R2 * w2) {
            const F wl = std::sqrt(w2);
            const V2T<F> uw = wv / wl;
            l.dir = V2T<F>{uw.y, -uw.x};
            u = (R * inv_th - wl) * uw;
        } else {
            const F leg = std::sqrt(d2 - R2);
            if (det(rp, wv) > F(0))
                l.dir = V2T<F>{rp.x * leg - rp.y * R, rp.x * R + rp.y * leg} / d2;
            else
                l.dir = -(V2T<F>{rp.x * leg + rp.y * R, -rp.x * R + rp.y * leg} / d2);
            const F dp2 = dot(rv, l.dir);
            u = dp2 * l.dir - rv;
        }
    } else {
        const F inv_dt = F(1) / time_step;
        const V2T<F> wv = rv - inv_dt * rp;
        const F wl = norm(wv);
        const V2T<F> uw = wv / wl;
        l.dir = V2T<F>{uw.y, -uw.x};
        u = (R * inv_dt - wl) * uw;
    }
    l.point = va + F(0.5) * u;
    return l;
}

using V2 = V2T<float>;
using HalfPlane = HalfPlaneT<float>;

} // namespace

extern "C" {

struct orc_orca_batch {
    int n_worlds, n_agents;
    /* ORCAParams, _orca.py:13-21 and PyRVOSimulator ctor :73-79 */
    float time_step, neighbor_dist;
    int max_neighbors;
    float time_horizon, radius;
    /* rvo2 agent state, single precision: (W,N,2) */
    float *pos, *vel, *pref;
    const float *max_speed;   /* (W,N) */
    const double *pref_speed; /* (W,N) adapter-side, fp64 (_orca.py:66,126) */
    double goal_threshold;    /* :67 */
    /* FixedWaypointTracker state (pedestrians/_fixed_waypoint_tracker.py) */
    int n_table;
    const double *table; /* (W,N,n_table,2) */
    int *wp_index;       /* (W,N) */
    double *cur_wp;      /* (W,N,2) */
    double reach;
    int loop;
    double *heading;     /* (W,N) atan2(vy,vx), _orca.py:103-104 */
    int *n_constraints;  /* (W,N) optional: ORCA half-planes used in the last step */
};

/* One RVO2 doStep for one world: new velocities for all agents from the OLD state, then update. */
static void orca_world_step(const orc_orca_batch *b, int w, std::vector<V2> &newv)
{
    const int n = b->n_agents;
    const V2 *pos = reinterpret_cast<const V2 *>(b->pos) + (size_t)w * n;
    const V2 *vel = reinterpret_cast<const V2 *>(b->vel) + (size_t)w * n;
    const V2 *pref = reinterpret_cast<const V2 *>(b->pref) + (size_t)w * n;
    const float *vmax = b->max_speed + (size_t)w * n;
    const int kmax = b->max_neighbors;
    std::vector<float> nd(kmax > 0 ? kmax : 1);
    std::vector<int> ni(kmax > 0 ? kmax : 1);
    std::vector<HalfPlane> hp(kmax > 0 ? kmax : 1);
    const float inv_th = 1.0f / b->time_horizon;
    for (int a = 0; a < n; ++a) {
        /* neighbour list: nearest `kmax` within neighbor_dist, insertion-sorted by squared distance */
        int cnt = 0;
        float range2 = b->neighbor_dist * b->neighbor_dist;
        if (kmax > 0) {
            for (int o = 0; o < n; ++o) {
                if (o == a) continue;
                const float d2 = norm2(pos[a] - pos[o]);
                if (!(d2 < range2)) continue;
                if (cnt < kmax) ++cnt;
                int i = cnt - 1;
                while (i != 0 && d2 < nd[i - 1]) {
                    nd[i] = nd[i - 1];
                    ni[i] = ni[i - 1];
                    --i;
                }
                nd[i] = d2;
                ni[i] = o;
                if (cnt == kmax) range2 = nd[cnt - 1];
            }
        }
        /* one ORCA half-plane per neighbour */
        for (int k = 0; k < cnt; ++k) {
            const int o = ni[k];
            hp[k] = orca_halfplane(pos[a], pos[o], vel[a], vel[o], b->radius, inv_th, b->time_step);
        }
        V2 res;
        const int fail = lp_plane(hp.data(), cnt, vmax[a], pref[a], false, res);
        if (fail < cnt) lp_relax(hp.data(), cnt, fail, vmax[a], res);
        newv[a] = res;
        if (b->n_constraints) b->n_constraints[(size_t)w * n + a] = cnt;
    }
}

/* _orca.py:118-127 */
static void orca_update_pref(orc_orca_batch *b, int w)
{
    const int n = b->n_agents;
    for (int i = 0; i < n; ++i) {
        const size_t gi = (size_t)w * n + i;
        const double dx = b->cur_wp[2 * gi] - (double)b->pos[2 * gi];
        const double dy = b->cur_wp[2 * gi + 1] - (double)b->pos[2 * gi + 1];
        const double nr = std::sqrt(dx * dx + dy * dy);
        if (nr <= b->goal_threshold) {
            b->pref[2 * gi] = 0.0f;
            b->pref[2 * gi + 1] = 0.0f;
        } else {
            b->pref[2 * gi] = (float)(dx / nr * b->pref_speed[gi]);
            b->pref[2 * gi + 1] = (float)(dy / nr * b->pref_speed[gi]);
        }
    }
}

void orc_orca_update_pref(orc_orca_batch *b)
{
    for (int w = 0; w < b->n_worlds; ++w) orca_update_pref(b, w);
}

/* ORCAPedestriansModel.step, _orca.py:95-112, n_steps times. */
int orc_orca_rollout(orc_orca_batch *b, int n_steps, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    std::atomic<int> next_world{0};
    auto worker = [&]() {
    for (int w = next_world++; w < b->n_worlds; w = next_world++) {
        const int n = b->n_agents;
        std::vector<V2> newv(n);
        for (int s = 0; s < n_steps; ++s) {
            orca_world_step(b, w, newv);
            for (int i = 0; i < n; ++i) {
                const size_t gi = (size_t)w * n + i;
                b->vel[2 * gi] = newv[i].x;
                b->vel[2 * gi + 1] = newv[i].y;
                b->pos[2 * gi] += newv[i].x * b->time_step;
                b->pos[2 * gi + 1] += newv[i].y * b->time_step;
                b->heading[gi] = std::atan2((double)newv[i].y, (double)newv[i].x);
            }
            /* tracker on the new positions (_orca.py:106-107); the steady flag is ignored */
            if (b->n_table > 0) {
                for (int i = 0; i < n; ++i) {
                    const size_t gi = (size_t)w * n + i;
                    const double dx = b->cur_wp[2 * gi] - (double)b->pos[2 * gi];
                    const double dy = b->cur_wp[2 * gi + 1] - (double)b->pos[2 * gi + 1];
                    if (!(std::sqrt(dx * dx + dy * dy) <= b->reach)) continue;
                    int ni = b->wp_index[gi] + 1;
                    if (ni >= b->n_table) ni = b->loop ? 0 : b->n_table - 1;
                    b->wp_index[gi] = ni;
                    b->cur_wp[2 * gi] = b->table[(gi * b->n_table + ni) * 2];
                    b->cur_wp[2 * gi + 1] = b->table[(gi * b->n_table + ni) * 2 + 1];
                }
            }
            orca_update_pref(b, w);
        }
    }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    return 0;
}


/* ---- test hooks (tests/test_oracle_orca_independent.py): the two building blocks on their own ---- */
/* out = {point.x, point.y, dir.x, dir.y} of the half-plane agent a takes on for agent o */
void orc_orca_halfplane(const float *pa, const float *po, const float *va, const float *vo, float radius, float time_horizon,
                        float time_step, float *out)
{
    const HalfPlane l = orca_halfplane(V2{pa[0], pa[1]}, V2{po[0], po[1]}, V2{va[0], va[1]}, V2{vo[0], vo[1]}, radius,
                                       1.0f / time_horizon, time_step);
    out[0] = l.point.x; out[1] = l.point.y; out[2] = l.dir.x; out[3] = l.dir.y;
}

/* new velocity for n half-planes lines = (n,4) {point, dir}, speed limit and preferred velocity; returns the index of the
 * first half-plane the 2-D program failed at (n = feasible; < n = the relaxed program ran) */
int orc_orca_solve(int n, const float *lines, float max_speed, const float *pref, float *out_v)
{
    std::vector<HalfPlane> hp(n > 0 ? n : 1);
    for (int k = 0; k < n; ++k) hp[k] = HalfPlane{V2{lines[4 * k], lines[4 * k + 1]}, V2{lines[4 * k + 2], lines[4 * k + 3]}};
    V2 res;
    const int fail = lp_plane(hp.data(), n, max_speed, V2{pref[0], pref[1]}, false, res);
    if (fail < n) lp_relax(hp.data(), n, fail, max_speed, res);
    out_v[0] = res.x; out_v[1] = res.y;
    return fail;
}

/* fp64 instantiation of the same algorithm -- NOT the oracle (rvo2 is single precision); test use only */
int orc_orca_solve_f64(int n, const double *lines, double max_speed, const double *pref, double *out_v)
{
    using H = HalfPlaneT<double>;
    using V = V2T<double>;
    std::vector<H> hp(n > 0 ? n : 1);
    for (int k = 0; k < n; ++k) hp[k] = H{V{lines[4 * k], lines[4 * k + 1]}, V{lines[4 * k + 2], lines[4 * k + 3]}};
    V res;
    const int fail = lp_plane(hp.data(), n, max_speed, V{pref[0], pref[1]}, false, res);
    if (fail < n) lp_relax(hp.data(), n, fail, max_speed, res);
    out_v[0] = res.x; out_v[1] = res.y;
    return fail;
}
void orc_orca_solve_batch_f64(int m, int n, const double *lines, const double *max_speed, const double *pref, double *out_v, int *out_fail)
{
    for (int q = 0; q < m; ++q)
        out_fail[q] = orc_orca_solve_f64(n, lines + (size_t)q * n * 4, max_speed[q], pref + 2 * q, out_v + 2 * q);
}
void orc_orca_halfplane_batch_f64(int m, const double *pa, const double *po, const double *va, const double *vo, double radius,
                                  double time_horizon, double time_step, double *out)
{
    using V = V2T<double>;
    for (int q = 0; q < m; ++q) {
        const HalfPlaneT<double> l = orca_halfplane(V{pa[2 * q], pa[2 * q + 1]}, V{po[2 * q], po[2 * q + 1]}, V{va[2 * q], va[2 * q + 1]},
                                                   V{vo[2 * q], vo[2 * q + 1]}, radius, 1.0 / time_horizon, time_step);
        out[4 * q] = l.point.x; out[4 * q + 1] = l.point.y; out[4 * q + 2] = l.dir.x; out[4 * q + 3] = l.dir.y;
    }
}

/* the same for m independent problems of n half-planes each (lines (m,n,4), max_speed (m), pref (m,2)) */
void orc_orca_solve_batch(int m, int n, const float *lines, const float *max_speed, const float *pref, float *out_v, int *out_fail)
{
    for (int q = 0; q < m; ++q)
        out_fail[q] = orc_orca_solve(n, lines + (size_t)q * n * 4, max_speed[q], pref + 2 * q, out_v + 2 * q);
}
void orc_orca_halfplane_batch(int m, const float *pa, const float *po, const float *va, const float *vo, float radius,
                              float time_horizon, float time_step, float *out)
{
    for (int q = 0; q < m; ++q)
        orc_orca_halfplane(pa + 2 * q, po + 2 * q, va + 2 * q, vo + 2 * q, radius, time_horizon, time_step, out + 4 * q);
}

} // extern "C"
