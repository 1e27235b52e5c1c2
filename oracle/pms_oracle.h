/*
 * pms_oracle.h -- CPU restatement (fp64, plain C) of the pyminisim pedestrian-dynamics hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package (pyminisim_b200/) may include, link,
 * load or call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the CPU baseline.
 *
 * Parity status: PINNED for HSFM, trackers, robot models, collision list, pedestrian detector,
 * lidar, omni detector -- against fixtures generated from the live reference
 * (tests/golden/make_golden.py, run in the build container where /root/reference is mounted).
 * ORCA lives in orca_oracle.cpp and is UNPINNED (third-party rvo2 is absent; see that file).
 *
 * Every function cites the reference file:line it restates (paths relative to the reference
 * repository root, TimeEscaper/pyminisim).
 */
#ifndef PMS_ORACLE_H
#define PMS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_ROBOT_NONE = 0, ORC_ROBOT_EXTERNAL = 1, ORC_ROBOT_UNICYCLE = 2, ORC_ROBOT_BICYCLE = 3 };
enum { ORC_TRACKER_FIXED = 0, ORC_TRACKER_QUEUE = 1 };
enum { ORC_MAP_EMPTY = 0, ORC_MAP_CIRCLES = 1, ORC_MAP_AABB = 2 };
enum { ORC_DET_POLAR = 0, ORC_DET_RELATIVE_ROTATED = 1, ORC_DET_RELATIVE = 2, ORC_DET_ABSOLUTE = 3 };

typedef struct orc_batch {
    int n_worlds, n_peds;
    /* HSFMParams (pedestrians/_hsfm.py:20-42) + ctor constants (:198-208,235-239) */
    double tau, A, B, k1, k2, ko, kd, k_lambda, alpha;
    double mass, ped_radius, hsfm_robot_radius;
    /* pedestrian state, (W,N,3) row-major: pose (x,y,theta), velocity (vx,vy,omega) */
    double *poses, *vels;
    double *vmag;               /* (W,N) desired speed */
    /* waypoint tracker */
    int tracker_kind;
    double reach;
    int loop;
    int n_table;                /* rows per pedestrian incl. the initial position at row 0 */
    const double *table;        /* (W,N,n_table,2) */
    int *wp_index;              /* (W,N) */
    double *cur_wp;             /* (W,N,2) current waypoint (kept in sync in both modes) */
    int q_cap;
    double *queue;              /* (W,N,q_cap,2) ring buffer */
    int *q_head, *q_count;      /* (W,N) */
    int ev_cap;
    int *ev_count;              /* scalar */
    int *events;                /* (ev_cap,3): step (0-based within this call), world, ped */
    /* robot */
    int robot_model;
    int robot_visible;
    double robot_radius;        /* physical radius: collision list + map veto */
    double wheel_base;
    double *robot_pose;         /* (W,3) */
    double *robot_vel;          /* (W,3) */
    double *robot_ctrl;         /* (W,2) */
    double *robot_rear;         /* (W,3) bicycle rear-axle pose */
    const double *controls;     /* (n_controls,W,2) or NULL: keep stored control */
    int n_controls;
    int *robot_world_collision; /* (W) */
    /* map */
    int map_kind, map_count, map_world_stride; /* stride in doubles between worlds; 0 = shared */
    const double *map_data;     /* circles (C,3) x,y,r  |  aabb (C,4) tl_x,tl_y,w,h */
    /* detector + collision list, evaluated after every step */
    int det_enabled;
    double det_max_dist, det_fov;
    int det_return_type;
    uint32_t *det_mask;         /* (W,ceil(N/32)) for the latest state */
    double *det_vals;           /* (W,N,2) */
    uint32_t *coll_mask;        /* (W,ceil(N/32)) */
    int64_t *det_count;         /* (W) accumulated detections over steps */
    int64_t *coll_count;        /* (W) accumulated robot-ped collisions over steps */
    int *nonfinite;             /* (W) 1-based step index (cumulative) of first non-finite state, 0 = finite */
    int64_t steps_done;         /* cumulative */
} orc_batch;

/* pedestrians/_hsfm.py:113-134 */
void orc_pair_force(double radius_i, double radius_j, const double *r_i, const double *r_j,
                    const double *v_i, const double *v_j, double A, double B, double k1, double k2,
                    double *f_out);

/* util/_util.py:5-6 and _hsfm.py:187-189 */
double orc_wrap_angle(double a);
double orc_pymod(double a, double b);

/* One HSFM step of one world, no tracker: pedestrians/_hsfm.py:46-110,146-158,250-299.
 * robot4 = {x,y,vx,vy} or NULL. poses/vels updated in place. */
void orc_hsfm_world_step(const orc_batch *p, int n, double *poses, double *vels, const double *wps,
                         const double *vmag, const double *robot4, double dt);

/* Full fused rollout: robot -> HSFM -> tracker/freeze -> collision list + detector; n_steps times. */
int orc_rollout(orc_batch *b, double dt, int n_steps, int n_threads);

/* sensors on the current state only (no stepping) */
void orc_sense(orc_batch *b, int n_threads);

/* sensors/_lidar.py:105-129 with world_map/_circles_world.py:41-50, _aabb_world.py:135-156.
 * hit_idx (W,n_beams) = sample index of first occupied sample or -1; points (W,n_beams,2). */
void orc_lidar_scan(const orc_batch *b, int n_beams, double max_dist, double resolution,
                    int32_t *hit_idx, double *points, int brute_force);

/* sensors/_omni_obstacle_detector.py:58-62 */
void orc_omni(const orc_batch *b, double *out);

/* world_map base is_collision: core/_world_map.py:35-62 */
int orc_map_is_collision(const orc_batch *b, int w, double x, double y, double radius);

#ifdef __cplusplus
}
#endif
#endif
