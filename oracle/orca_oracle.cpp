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

constexpr float kEps = 1e-5f;

struct V2 {
    float x, y;
};
inline V2 operator+(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
inline V2 operator-(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
inline V2 operator-(V2 a) { return {-a.x, -a.y}; }
inline V2 operator*(float s, V2 a) { return {s * a.x, s * a.y}; }
inline V2 operator/(V2 a, float s) {
    const float inv = 1.0f / s; /* the library divides through a reciprocal */
    return {a.x * inv, a.y * inv};
}
inline float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
inline float det(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
inline float norm2(V2 a) { return dot(a, a); }
inline float norm(V2 a) { return std::sqrt(norm2(a)); }
inline V2 unit(V2 a) { return a / norm(a); }

struct HalfPlane {
    V2 point, dir;
};

/* 1-D LP along half-plane `k`, constrained by the speed disc and half-planes 0..k-1. */
bool lp_line(const HalfPlane *hp, int k, float radius, V2 opt, bool direction_opt, V2 &result)
{
    const float dp = dot(hp[k].point, hp[k].dir);
    const float disc = dp * dp + radius * radius - norm2(hp[k].point);
    if (disc < 0.0f) return false;
    const float sq = std::sqrt(disc);
    float t_lo = -dp - sq, t_hi = -dp + sq;
    for (int i = 0; i < k; ++i) {
        const float den = det(hp[k].dir, hp[i].dir);
        const float num = det(hp[i].dir, hp[k].point - hp[i].point);
        if (std::fabs(den) <= kEps) {
            if (num < 0.0f) return false;
            continue;
        }
        const float t = num / den;
        if (den >= 0.0f) t_hi = std::fmin(t_hi, t);
        else t_lo = std::fmax(t_lo, t);
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

/* 2-D incremental LP. Returns the index of the first half-plane that made it infeasible, or n. */
int lp_plane(const HalfPlane *hp, int n, float radius, V2 opt, bool direction_opt, V2 &result)
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

/* Fallback: minimise the maximum penetration (no obstacle half-planes on this path). */
void lp_relax(const HalfPlane *hp, int n, int begin, float radius, V2 &result)
{
    float distance = 0.0f;
    std::vector<HalfPlane> proj;
    for (int i = begin; i < n; ++i) {
        if (det(hp[i].dir, hp[i].point - result) > distance) {
            proj.clear();
            for (int j = 0; j < i; ++j) {
                HalfPlane l;
                const float d = det(hp[i].dir, hp[j].dir);
                if (std::fabs(d) <= kEps) {
                    if (dot(hp[i].dir, hp[j].dir) > 0.0f) continue;
                    l.point = 0.5f * (hp[i].point + hp[j].point);
                } else {
                    l.point = hp[i].point + (det(hp[j].dir, hp[i].point - hp[j].point) / d) * hp[i].dir;
                }
                l.dir = unit(hp[j].dir - hp[i].dir);
                proj.push_back(l);
            }
            const V2 keep = result;
            if (lp_plane(proj.data(), (int)proj.size(), radius, V2{-hp[i].dir.y, hp[i].dir.x}, true, result) <
                (int)proj.size())
                result = keep;
            distance = det(hp[i].dir, hp[i].point - result);
        }
    }
}

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
            const V2 rp = pos[o] - pos[a];
            const V2 rv = vel[a] - vel[o];
            const float d2 = norm2(rp);
            const float R = b->radius + b->radius;
            const float R2 = R * R;
            V2 u;
            HalfPlane l;
            if (d2 > R2) {
                const V2 wv = rv - inv_th * rp;
                const float w2 = norm2(wv);
                const float dp1 = dot(wv, rp);
                if (dp1 < 0.0f && dp1 * dp1 > R2 * w2) {
                    const float wl = std::sqrt(w2);
                    const V2 uw = wv / wl;
                    l.dir = V2{uw.y, -uw.x};
                    u = (R * inv_th - wl) * uw;
                } else {
                    const float leg = std::sqrt(d2 - R2);
                    if (det(rp, wv) > 0.0f)
                        l.dir = V2{rp.x * leg - rp.y * R, rp.x * R + rp.y * leg} / d2;
                    else
                        l.dir = -(V2{rp.x * leg + rp.y * R, -rp.x * R + rp.y * leg} / d2);
                    const float dp2 = dot(rv, l.dir);
                    u = dp2 * l.dir - rv;
                }
            } else {
                const float inv_dt = 1.0f / b->time_step;
                const V2 wv = rv - inv_dt * rp;
                const float wl = norm(wv);
                const V2 uw = wv / wl;
                l.dir = V2{uw.y, -uw.x};
                u = (R * inv_dt - wl) * uw;
            }
            l.point = vel[a] + 0.5f * u;
            hp[k] = l;
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

} // extern "C"
