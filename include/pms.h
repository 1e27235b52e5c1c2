/*
 * pms.h -- C ABI of libpms_b200: the B200-native (sm_100a) pedestrian-dynamics hot path of
 * TimeEscaper/pyminisim, batched over independent worlds.
 *
 * The reference has no FFI of its own: its boundary for this path is the duck-typed Python
 * interface  AbstractPedestriansModel.step(dt, robot_pose, robot_velocity)
 * (pyminisim/core/_pedestrians_model.py:34-47), AbstractSensor.get_reading(world_state, world_map)
 * (core/_sensor.py:16-38) and Simulation.step / _get_world_state / _update_sensors_state
 * (core/_simulation.py:60-73,138-185).  Each entry point below names the reference code it
 * replaces.  INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions: plain C, opaque handle, every call returns 0 on success or a negative pms_status
 * (message via pms_last_error()).  All host arrays are caller-owned, row-major, float64 unless
 * stated, and world-major: "(W,N,3)" is poses[w][i][0..2].  One handle per device; a handle is
 * not thread-safe (one host thread per GPU).  No exceptions cross the boundary.
 */
#ifndef PMS_H
#define PMS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pms_sim pms_sim;

typedef enum {
    PMS_OK = 0,
    PMS_ERR_INVALID = -1,   /* bad argument (AssertionError / ValueError in the reference) */
    PMS_ERR_CUDA = -2,      /* CUDA runtime failure */
    PMS_ERR_STATE = -3,     /* call order / configuration error (RuntimeError in the reference) */
    PMS_ERR_UNSUPPORTED = -4
} pms_status;

/* Arithmetic mode of the HSFM kernels.
 *   PMS_FP32  : pair forces and per-pedestrian math in fp32 (north-star production mode).  The state is carried as
 *               unevaluated fp32 sums hi + lo with compensated Euler updates, so a step no longer rounds the full
 *               magnitude of the state to fp32 (the error that dominates plain fp32 and that the crowd dynamics
 *               amplify by orders of magnitude over 1000 steps); pair forces see the high parts.
 *   PMS_FP32_PLAIN : as PMS_FP32 with a plain fp32 state, one rounding per variable per step.
 *   PMS_MIXED : fp64 state and per-pedestrian math; pair forces in fp32 from hi/lo-split
 *               positions (differences exact to fp64).
 *   PMS_FP64  : everything fp64 -- verification mode, algorithmically identical to the reference.
 * Robot pose, detector and collision predicates are evaluated in fp64 in every mode. */
typedef enum { PMS_FP32 = 0, PMS_MIXED = 1, PMS_FP64 = 2, PMS_FP32_PLAIN = 3 } pms_precision;

typedef enum { PMS_ROBOT_NONE = 0, PMS_ROBOT_EXTERNAL = 1, PMS_ROBOT_UNICYCLE = 2, PMS_ROBOT_BICYCLE = 3 } pms_robot_model;
typedef enum { PMS_MAP_EMPTY = 0, PMS_MAP_CIRCLES = 1, PMS_MAP_AABB = 2 } pms_map_kind;
typedef enum { PMS_DET_POLAR = 0, PMS_DET_RELATIVE_ROTATED = 1, PMS_DET_RELATIVE = 2, PMS_DET_ABSOLUTE = 3 } pms_det_return;

/* HSFMParams (pedestrians/_hsfm.py:20-42) + constructor constants (:198-208,235-239). */
typedef struct {
    double tau, A, B, k_1, k_2, k_o, k_d, k_lambda, alpha;
    double pedestrian_mass;   /* 70.0 */
    double pedestrian_radius; /* core/_constants.py:2 = 0.3 */
    double robot_radius;      /* HSFM ctor robot_radius = 0.35: used in the robot force term only */
} pms_hsfm_params;

/* ORCAParams (pedestrians/_orca.py:13-21) + PyRVOSimulator ctor (:73-79). */
typedef struct {
    double time_step, neighbor_dist;
    int max_neighbors;
    double time_horizon, radius, goal_reaching_threshold;
} pms_orca_params;

const char *pms_last_error(void);
int pms_version(void);
void pms_default_hsfm_params(pms_hsfm_params *out);   /* HSFMParams.create_default(), _hsfm.py:31-42 */
void pms_default_orca_params(pms_orca_params *out);   /* ORCAParams(), _orca.py:13-21; time_step 0.01 */

/* ---- lifetime ---- */
/* Replaces HeadedSocialForceModelPolicy.__init__ (_hsfm.py:198-244) for n_worlds independent
 * simulations of n_peds pedestrians each.  Fails (PMS_ERR_CUDA) when no CUDA device is present:
 * there is no CPU fallback. */
int pms_create(int device, int n_worlds, int n_peds, int precision, const pms_hsfm_params *params, pms_sim **out);
int pms_destroy(pms_sim *h);
int pms_n_worlds(const pms_sim *h);
int pms_n_peds(const pms_sim *h);

/* ---- pedestrian state: AbstractPedestriansModelState.poses / .velocities
 *      (core/_pedestrians_model.py:9-31) and reset_to_state (_hsfm.py:308-310) ---- */
int pms_set_ped_state(pms_sim *h, const double *poses /*(W,N,3)*/, const double *vels /*(W,N,3)*/);
int pms_get_ped_state(pms_sim *h, double *poses /*(W,N,3)*/, double *vels /*(W,N,3)*/);
/* The same with float32 host arrays: half the PCIe bytes for callers that do not need more than the FP32 kernels compute in
 * (the state of PMS_FP32 is hi + lo: the getter returns the correctly rounded sum, the setter starts from lo = 0). */
int pms_set_ped_state_f32(pms_sim *h, const float *poses /*(W,N,3)*/, const float *vels /*(W,N,3) or NULL = 0*/);
int pms_get_ped_state_f32(pms_sim *h, float *poses /*(W,N,3) or NULL*/, float *vels /*(W,N,3) or NULL*/);
/* Device-resident I/O: `poses` / `vels` are DEVICE pointers to caller-owned (W,N,3) arrays of float64 (is_f32 = 0) or
 * float32 (is_f32 = 1) on the handle's device -- e.g. the data_ptr() of a torch tensor.  One conversion kernel on the handle's
 * stream, no host synchronisation and no PCIe traffic: order the consumer after pms_sync or after the handle's stream. */
int pms_set_ped_state_device(pms_sim *h, const void *poses, const void *vels /*or NULL = 0*/, int is_f32);
int pms_get_ped_state_device(pms_sim *h, void *poses /*or NULL*/, void *vels /*or NULL*/, int is_f32);
/* Raw device pointers of the handle's result buffers (zero-copy for callers that stay on the GPU): valid until the handle is
 * destroyed or -- the PMS_BUF_REC_* record buffers -- until the next launch grows them; written by the launches on the
 * handle's stream.  bytes = current extent (records: the readings of the last launch).  *ptr = NULL when the buffer does not
 * exist (kind not scheduled). */
typedef enum {
    PMS_BUF_DET_MASK = 0,        /* uint32 (W, ceil(N/32)): detector membership after the last step */
    PMS_BUF_COLL_MASK = 1,       /* uint32 (W, ceil(N/32)): robot-pedestrian collision list after the last step */
    PMS_BUF_DET_VALUES = 2,      /* float64 (W, N, 2) */
    PMS_BUF_ROBOT_POSE = 3,      /* float64 (W, 3) */
    PMS_BUF_ROBOT_VELOCITY = 4,  /* float64 (W, 3) */
    PMS_BUF_DETECTIONS = 5,      /* uint64 (W): episode counter */
    PMS_BUF_COLLISIONS = 6,      /* uint64 (W): episode counter */
    PMS_BUF_FIRST_NONFINITE = 7, /* int32 (W): 1-based step of the first non-finite state, INT32_MAX = never */
    PMS_BUF_WAYPOINT_INDEX = 8,  /* int32 (W, N) */
    PMS_BUF_REC_DET_MASK = 16, PMS_BUF_REC_DET_VALUES = 17, PMS_BUF_REC_COLL_MASK = 18, PMS_BUF_REC_LIDAR_HIT = 19,
    PMS_BUF_REC_LIDAR_POINTS = 20, PMS_BUF_REC_OMNI = 21   /* the records of pms_sensor_schedule, layouts as in pms_get_*_records */
} pms_buffer;
int pms_device_buffer(pms_sim *h, int which, void **ptr, uint64_t *bytes /*or NULL*/);
int pms_set_ped_speeds(pms_sim *h, const double *vmag /*(W,N)*/); /* pedestrian_linear_velocity_magnitude */
/* Asynchronous I/O mode (pipelining several handles): pms_set_ped_state, pms_get_ped_state, pms_get_detector,
 * pms_get_collisions and pms_get_stats enqueue their copies on the handle's stream and return; the host buffers must be
 * page-locked (pms_host_alloc) and stay valid until pms_sync.  pms_get_stats then leaves INT32_MAX ("never") in
 * first_nonfinite_step instead of 0. */
int pms_set_async_io(pms_sim *h, int enabled);
/* device-side snapshot of the complete simulation state (checkpoint / resume) */
/* (pms_set_waypoints_fixed and pms_set_waypoint_queue replace device tables a snapshot refers to: they discard a saved
 * snapshot, pms_snapshot_restore then fails with PMS_ERR_STATE until the next pms_snapshot_save) */
int pms_snapshot_save(pms_sim *h);
int pms_snapshot_restore(pms_sim *h);

/* ---- waypoint trackers ---- */
/* FixedWaypointTracker (pedestrians/_fixed_waypoint_tracker.py:28-98).  table holds n_rows rows
 * per pedestrian with the initial position at row 0 (:45); start_index NULL -> 1 (:59). */
int pms_set_waypoints_fixed(pms_sim *h, const double *table /*(W,N,n_rows,2)*/, int n_rows,
                            const int32_t *start_index /*(W,N) or NULL*/, double reach_distance, int loop);
/* RandomWaypointTracker (pedestrians/_random_waypoint_tracker.py:66-91): the device pops
 * current <- queue front on reach and logs an event; the host draws the new tail with the global
 * numpy RNG while replaying the event log in order, then appends it. */
int pms_set_waypoint_queue(pms_sim *h, const double *current /*(W,N,2)*/, const double *queue /*(W,N,q_len,2)*/,
                           int q_len, double reach_distance);
int pms_append_waypoints(pms_sim *h, int n, const int32_t *world_ped /*(n,2)*/, const double *points /*(n,2)*/);
int pms_pop_waypoint_events(pms_sim *h, int32_t *events /*(cap,3): step,world,ped*/, int cap, int *n_out);
int pms_get_waypoints(pms_sim *h, double *current /*(W,N,2)*/, int32_t *index /*(W,N) or NULL*/);

/* ---- robot: AbstractRobotMotionModel (core/_motion.py:10-32) ---- */
/* model EXTERNAL: pose/velocity supplied by the caller are what AbstractPedestriansModel.step
 * receives (held for all steps of the next pms_hsfm_step call).  UNICYCLE / BICYCLE integrate on
 * the device exactly like robot/_unicycle.py:53-69 and robot/_bicycle.py:71-95 including the map
 * collision veto (core/_world_map.py:35-62).  `radius` is AbstractRobotMotionModel.radius
 * (collision list, veto); `visible` is HSFM's robot_visible (_hsfm.py:279). */
int pms_set_robot(pms_sim *h, int model, const double *pose /*(W,3)*/, const double *velocity /*(W,3) EXTERNAL only, else NULL*/,
                  const double *control /*(W,2) or NULL*/, double radius, double wheel_base, int visible);
/* controls for the next pms_hsfm_step call: row s drives step s; the last row is held. NULL/0 clears. */
int pms_set_robot_controls(pms_sim *h, const double *controls /*(n_rows,W,2)*/, int n_rows);
/* the same from a DEVICE array (n_rows,W,2) of float64 owned by the caller (e.g. a policy's output tensor): read in place by
 * the following launches -- keep it alive and complete until they have run; NULL/0 clears */
int pms_set_robot_controls_device(pms_sim *h, const double *controls_dev, int n_rows);
int pms_get_robot(pms_sim *h, double *pose /*(W,3)*/, double *velocity /*(W,3)*/, int32_t *world_collision /*(W) or NULL*/);

/* ---- world map: world_map/_circles_world.py, _aabb_world.py, _empty_world.py ---- */
int pms_set_map(pms_sim *h, int kind, const double *data /*circles (C,3) | aabb (C,4); per-world: (W,C,.)*/,
                int count, int per_world);

/* HSFM acceleration noise (noise_std, _hsfm.py:288-289): `noise` (W,N,2) is added to the body-frame
 * acceleration dv_b in every step of the following pms_hsfm_step calls; the caller draws it (the
 * reference uses the global numpy RNG).  NULL switches it off. */
int pms_set_accel_noise(pms_sim *h, const double *noise /*(W,N,2) or NULL*/);
/* The same noise drawn on the device: N(0, noise_std) per pedestrian and component, afresh in EVERY step of a fused
 * multi-step launch like np.random.normal in _hsfm.py:289 (the array above is constant over a launch).  Counter-based
 * Philox4x32-10 keyed by (seed, world*N + pedestrian, absolute step): the stream does not depend on the launch
 * configuration, the kernel chosen or the split of an episode into calls.  Statistical parity with the reference (its
 * values come from the global numpy Mersenne twister), not bit parity.  noise_std = 0 switches it off. */
int pms_set_accel_noise_std(pms_sim *h, double noise_std, uint64_t seed);
/* The generator itself, on the host, so that a caller can reproduce a device stream: Philox4x32-10 (Salmon et al. 2011;
 * Random123 known-answer vectors in tests/test_cabi_symbols.py).  The device draws block
 *   ctr = {id lo, id hi, step lo, step hi ^ (stream << 28)}, key = {seed lo, seed hi},
 * id = world * n_peds + pedestrian, stream 0 = acceleration noise (words 0, 1 -> Box-Muller pair), 1 = detector noise
 * (word 0 -> misdetection uniform, words 1, 2 -> Box-Muller pair for distance / angle); uniform = (word + 0.5) / 2^32. */
void pms_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* LidarSensorNoise (sensors/_lidar.py:52-87) applied by pms_lidar_scan: hit points dropped with probability drop_prob
 * (hit index -1), N(point_mean, point_cov) added to the survivors; point_mean (2) / point_cov (2,2) NULL = the reference's
 * defaults (0, diag(0.001)).  Stream 2, id = world * n_beams + beam.  OmniObstacleDetectorNoise
 * (sensors/_omni_obstacle_detector.py:25-37,64-69) applied by pms_omni_scan: inf with probability misdetection_prob,
 * else distance + N(mu, sigma).  Stream 3, id = world.  Statistical parity. */
int pms_set_lidar_noise(pms_sim *h, int enabled, double drop_prob, const double *point_mean, const double *point_cov, uint64_t seed);
int pms_set_omni_noise(pms_sim *h, int enabled, double misdetection_prob, double distance_mu, double distance_sigma, uint64_t seed);

/* ---- sensors ---- */
/* PedestrianDetector (sensors/_pedestrian_detector.py:78-134); fov in radians. */
int pms_detector_config(pms_sim *h, int enabled, double max_dist, double fov, int return_type);
/* PedestrianDetectorNoise (sensors/_pedestrian_detector.py:60-72,104-112) on the device: pms_get_detector drops every
 * detected pedestrian with probability misdetection_prob, adds N(distance_mu, distance_sigma) / N(angle_mu, angle_sigma)
 * to the survivors in the polar frame and converts to the configured return type (:114-134).  Philox stream keyed by
 * (seed, pedestrian, absolute step): repeated reads of one step return the same noisy reading.  Statistical parity. */
int pms_set_detector_noise(pms_sim *h, int enabled, double misdetection_prob, double distance_mu, double distance_sigma,
                           double angle_mu, double angle_sigma, uint64_t seed);

/* ---- the hot path ---- */
/* n_steps fused iterations of Simulation._make_steps + _get_world_state + detector
 * (core/_simulation.py:119-160; pedestrians/_hsfm.py:250-306): robot step -> HSFM forces and
 * explicit-Euler integration -> waypoint hand-off / freeze -> collision list + detector. */
int pms_hsfm_step(pms_sim *h, double dt, int n_steps);
/* same, asynchronous on the handle's stream (no host synchronisation); pair with pms_sync */
int pms_hsfm_step_async(pms_sim *h, double dt, int n_steps);
int pms_sync(pms_sim *h);
/* One tick of the reference's step loop in one call and one host synchronisation -- what Simulation.step does between
 * two readings (core/_simulation.py:60-73,119-160): AbstractPedestriansModel.step(dt, robot_pose, robot_velocity)
 * (pedestrians/_hsfm.py:250-306) with the robot supplied by the caller (model EXTERNAL; NULL pose = no robot,
 * _hsfm.py:279-284), then the state, the waypoint indices, the robot-pedestrian collision list, the detector reading, the
 * non-finite flag and the number of logged waypoint events.  Equivalent to pms_set_robot + pms_hsfm_step +
 * pms_get_ped_state + pms_get_waypoints + pms_get_collisions + pms_get_detector + pms_get_stats; every output pointer may
 * be NULL.  `robot_radius` is AbstractRobotMotionModel.radius (collision list). */
int pms_step_tick(pms_sim *h, double dt, int n_steps, const double *robot_pose /*(W,3) or NULL*/,
                  const double *robot_velocity /*(W,3) or NULL*/, double robot_radius, int visible,
                  double *poses /*(W,N,3)*/, double *vels /*(W,N,3)*/, int32_t *wp_index /*(W,N)*/,
                  uint32_t *coll_mask /*(W,ceil(N/32))*/, uint32_t *det_mask /*(W,ceil(N/32))*/, double *det_vals /*(W,N,2)*/,
                  int32_t *first_nonfinite_step /*(W)*/, int32_t *n_events /*(1)*/);
/* The sensor pass of one tick for pedestrians that were moved elsewhere (ORCA handle, ReplayPedestriansPolicy, a custom
 * model): Simulation._get_world_state's collision list (core/_simulation.py:138-160) and PedestrianDetector.get_reading
 * (sensors/_pedestrian_detector.py:78-134) for the given poses and robot pose, in one call and one host synchronisation.
 * Equivalent to pms_set_ped_state + pms_set_robot(EXTERNAL) + pms_sense + pms_get_collisions + pms_get_detector. */
int pms_sense_tick(pms_sim *h, const double *poses /*(W,N,3)*/, const double *robot_pose /*(W,3)*/, double robot_radius,
                   uint32_t *coll_mask /*(W,ceil(N/32)) or NULL*/, uint32_t *det_mask /*or NULL*/, double *det_vals /*(W,N,2) or NULL*/);
/* CUDA stream of the handle as an opaque pointer (cudaStream_t), for event timing by the caller */
void *pms_stream(pms_sim *h);
/* evaluate collision list + detector on the current state without stepping (Simulation.__init__
 * runs the sensor pass once, core/_simulation.py:37) */
int pms_sense(pms_sim *h);

/* readings for the latest state */
int pms_get_detector(pms_sim *h, uint32_t *mask /*(W,ceil(N/32))*/, double *values /*(W,N,2)*/);
int pms_get_detector_f32(pms_sim *h, uint32_t *mask /*(W,ceil(N/32))*/, float *values /*(W,N,2)*/);   /* exact reading, float32 values */
int pms_get_collisions(pms_sim *h, uint32_t *mask /*(W,ceil(N/32))*/);
/* per-world episode statistics accumulated over all steps since the last pms_reset_stats */
int pms_get_stats(pms_sim *h, int64_t *detections /*(W)*/, int64_t *collisions /*(W)*/, int32_t *first_nonfinite_step /*(W)*/);
int pms_reset_stats(pms_sim *h);

/* LidarSensor (sensors/_lidar.py:105-129) against the map; hit_index = index of the first
 * occupied sample along the beam or -1; points are absolute. */
int pms_lidar_scan(pms_sim *h, int n_beams, double max_dist, double resolution,
                   int32_t *hit_index /*(W,n_beams)*/, double *points /*(W,n_beams,2)*/);
/* OmniObstacleDetector (sensors/_omni_obstacle_detector.py:58-62): min distance robot -> map */
int pms_omni_scan(pms_sim *h, double *min_dist /*(W)*/);

/* ---- sensors inside a fused launch: the hold-time scheduler of Simulation._update_sensors_state
 *      (core/_simulation.py:162-185) on the device ----
 * The reference keeps one hold_time per sensor: every step hold += dt, and when hold >= period the sensor is read and
 * hold = 0 (dt = 0.01: a 10 Hz sensor fires every ELEVEN steps).  A sensor kind registered here is scheduled the same way
 * inside pms_hsfm_step / pms_hsfm_step_async / pms_step_tick: a tiny kernel replays the fp64 recurrence for the n_steps of
 * the launch, the step kernel takes the reading after every firing step and appends it to the RECORDS of the launch --
 * record k of a kind is its k-th reading of the LAST launch, taken after launch step steps[k] (0-based).  hold_time
 * persists from launch to launch (and through pms_snapshot_save / restore), so an episode split into several launches
 * fires at the same steps as one long launch.  The kinds:
 *   PMS_SENSOR_DETECTOR   PedestrianDetector.get_reading (sensors/_pedestrian_detector.py:78-134): membership mask per
 *                         record, plus the (W,N,2) values when flags has PMS_SENSOR_VALUES (exact, no device noise)
 *   PMS_SENSOR_LIDAR      LidarSensor.get_reading (sensors/_lidar.py:105-129) with the pms_lidar_config geometry, taken by
 *                         the warp that integrates the robot; LidarSensorNoise as set by pms_set_lidar_noise
 *   PMS_SENSOR_OMNI       OmniObstacleDetector.get_reading (sensors/_omni_obstacle_detector.py:58-62), pms_set_omni_noise
 *   PMS_SENSOR_COLLISIONS WorldState.robot_to_pedestrians_collisions (core/_simulation.py:141-143); not a sensor in the
 *                         reference (every step carries it): schedule it with period 0 for a per-step history
 * period = AbstractSensor.period (1 / frequency; 0 = a reading after every step); hold_time = SensorState.hold_time to
 * start from (0 right after Simulation.__init__, whose own first reading is pms_sense / pms_lidar_scan / pms_omni_scan).
 * Without a robot (PMS_ROBOT_NONE) the detector / collision records are empty masks and lidar / omni records are not
 * written (the reference returns no sensor state at all, core/_simulation.py:164-165). */
typedef enum { PMS_SENSOR_DETECTOR = 0, PMS_SENSOR_LIDAR = 1, PMS_SENSOR_OMNI = 2, PMS_SENSOR_COLLISIONS = 3 } pms_sensor_kind;
#define PMS_SENSOR_VALUES 1
int pms_lidar_config(pms_sim *h, int n_beams, double max_dist, double resolution);   /* LidarSensorConfig, _lidar.py:8-23 */
int pms_sensor_schedule(pms_sim *h, int kind, int enabled, double period, double hold_time, int flags);
/* the readings of the last launch: how many, after which launch steps, and the hold_time the scheduler is left with */
int pms_get_sensor_fires(pms_sim *h, int kind, int32_t *steps /*(cap) or NULL*/, int cap, int *n_fires, double *hold_time /*or NULL*/);
/* records first .. first + count - 1 of the last launch (host arrays, record-major) */
int pms_get_detector_records(pms_sim *h, int first, int count, uint32_t *mask /*(count,W,ceil(N/32))*/, double *values /*(count,W,N,2) or NULL*/);
int pms_get_collision_records(pms_sim *h, int first, int count, uint32_t *mask /*(count,W,ceil(N/32))*/);
int pms_get_lidar_records(pms_sim *h, int first, int count, int32_t *hit_index /*(count,W,n_beams)*/, double *points /*(count,W,n_beams,2) or NULL*/);
int pms_get_omni_records(pms_sim *h, int first, int count, double *min_dist /*(count,W)*/);

/* ---- recorded trajectories: ReplayPedestriansPolicy (pedestrians/_replay.py:19-59) as a batched gather ----
 * The frames (T,W,N,3) are uploaded once and stay in HBM.  pms_replay_step plays n_steps of them in ONE launch: per step
 * the robot model steps (pms_set_robot; device models or the external pose), the frame counter advances by the rule of
 * _replay.py:46-52 (next frame; at the last frame: back to 0 when loop, else stay), and the collision list + pedestrian
 * detector are evaluated on the frame's poses (exact fp64 predicates) -- with the records of pms_sensor_schedule when kinds
 * are scheduled.  The frame shown becomes the pedestrian state (pms_get_ped_state).  pms_replay_frame jumps
 * (reset_to_state, _replay.py:57-59: set_frame >= 0) or queries (-1).  pms_replay_tick = one tick of the drop-in class in
 * one call and one synchronisation: external robot pose in (NULL = no robot), collision list + detector reading out. */
int pms_set_replay(pms_sim *h, const double *poses /*(T,W,N,3)*/, const double *vels /*(T,W,N,3)*/, int n_frames, int loop);
int pms_replay_step(pms_sim *h, double dt, int n_steps);
int pms_replay_frame(pms_sim *h, int set_frame, int *frame_out /*or NULL*/);
int pms_replay_tick(pms_sim *h, int n_steps, const double *robot_pose /*(W,3) or NULL*/, double robot_radius,
                    uint32_t *coll_mask /*(W,ceil(N/32)) or NULL*/, uint32_t *det_mask /*or NULL*/, double *det_vals /*(W,N,2) or NULL*/,
                    int32_t *frame_out /*or NULL*/);

/* ---- ORCA (pedestrians/_orca.py:95-127 around rvo2.doStep) ---- */
int pms_orca_create(int device, int n_worlds, int n_agents, const pms_orca_params *params, pms_sim **out);
int pms_orca_set_agents(pms_sim *h, const double *positions /*(W,N,2)*/, const double *velocities /*(W,N,2)*/,
                        const double *max_speeds /*(W,N)*/, const double *pref_speeds /*(W,N)*/);
/* FixedWaypointTracker for the ORCA agents (same semantics as pms_set_waypoints_fixed; the steady
 * flag is ignored like _orca.py:106-107 does) */
int pms_orca_set_waypoints_fixed(pms_sim *h, const double *table /*(W,N,n_rows,2)*/, int n_rows,
                                 const int32_t *start_index /*(W,N) or NULL*/, double reach_distance, int loop);
/* Any other tracker (AbstractWaypointTracker, e.g. RandomWaypointTracker: _orca.py:106-107 accepts them all): the tracker runs
 * on the host, the caller hands over tracker.state.current_waypoints after every update; the device hand-off is off and
 * the preferred velocities (_orca.py:118-127) are recomputed from the given waypoints and the current positions. */
int pms_orca_set_waypoints_current(pms_sim *h, const double *current /*(W,N,2)*/);
/* enabled = 0: the waypoint setters above leave the agents' preferred velocities untouched -- the semantics of
 * ORCAPedestriansModel.reset_to_state (_orca.py:114-116), which resets the tracker only: the next doStep still runs on the old
 * preferred velocities and _update_preferred_velocities (:118-127) follows after it.  Default 1 (constructor semantics, :86). */
int pms_orca_set_pref_refresh(pms_sim *h, int enabled);
int pms_orca_step(pms_sim *h, int n_steps);
int pms_orca_step_async(pms_sim *h, int n_steps);          /* no host synchronisation; pair with pms_sync */
/* one tick of ORCAPedestriansModel.step (_orca.py:95-112): n_steps doStep and the agents' poses / velocities back in one call */
int pms_orca_tick(pms_sim *h, int n_steps, double *poses /*(W,N,3)*/, double *velocities /*(W,N,3)*/);
/* device-side snapshot of the agents (positions, velocities, preferred velocities, waypoint indices): checkpoint / resume */
int pms_orca_snapshot_save(pms_sim *h);
int pms_orca_snapshot_restore(pms_sim *h);
int pms_orca_get_agents(pms_sim *h, double *poses /*(W,N,3)*/, double *velocities /*(W,N,3)*/);

/* ---- launch introspection (bench / tests) ---- */
typedef struct {
    int kernel;            /* 0 = directed-pair small-world kernel, 1 = large-crowd kernel, 2 = ORCA, 3 = warp-per-world kernel,
                              4 = pair-warp / owner-warp kernel (under-filled batches, N <= 64), 5 = replay kernel */
    int threads_per_block, blocks, worlds_per_block, j_split; /* large kernel: j_split = cluster size */
    int smem_bytes;
    int n_chunks;          /* > 1: persistent CTAs pull (chunk, world-group) items dynamically */
    int blocks_per_sm;     /* resident CTAs per SM (occupancy API) */
    int64_t kernel_launches; /* cumulative count of kernels launched by this handle */
} pms_launch_info;
int pms_get_launch_info(pms_sim *h, pms_launch_info *out);
/* tuning overrides; 0 = automatic.
 *   j_split        : FP32 / MIXED worlds of up to 128 pedestrians run the warp-per-world symmetric-pair kernel; there
 *                    j_split 1..3 selects a smaller register budget (96 / 80 / 80 registers per thread, more resident warps), and a negative
 *                    value selects the directed-pair kernel instead (-1: automatic j-slices, -1-S: S slices).  For the
 *                    directed-pair kernel (FP64, larger worlds) j_split = 1, 2, 4, 8 is the number of j-slices.
 *                    FP32 batches of up to 64 pedestrians per world that fit about one wave of CTAs (one world per CTA) run the
 *                    pair-warp / owner-warp kernel instead (kernel 4: 3T + 1 warps per world, the pair pass of step t + 1
 *                    overlaps the dynamics of step t); j_split 4 forces it (worlds_per_block 1 .. 5 then also forces the
 *                    variant for two-tile worlds -- one / two units of the pair scheme per pair warp / two units compiled for 7,
 *                    4, 3 worlds per SM: bit-identical results, different speed; pms_launch_info.j_split reports the variant),
 *                    j_split 5 forces the warp-per-world kernel, and so
 *                    does any explicit worlds_per_block / n_chunks.  The two kernels add the same pair forces in a different
 *                    order: results agree to fp32 rounding, not bit for bit -- pin one of them (4 / 5) when the bits of a world
 *                    must not depend on the size of the batch it is in.
 *   n_chunks       : 1 forces the static grid; for the large-crowd kernel -2 keeps the caller's pedestrian order (no
 *                    sorted working copy) and -1 also visits every j-tile (no exact pruning).
 *   cluster_size   : (1,2,4,8) multicast width of the large-crowd kernel; a negative value also forces that kernel
 *                    (-100: forces the large-crowd path with its default kernel choice)
 *                    for small N. */
int pms_set_tuning(pms_sim *h, int j_split, int worlds_per_block, int n_chunks, int cluster_size);

/* measured FP32 FMA-pipe peak (TFLOP/s, FMA = 2 FLOP) of `device`: the roofline denominator of the
 * pairwise kernel (MEASURED_PEAKS.json holds only HBM and tensor-core peaks) */
double pms_measure_fp32_tflops(int device);

/* pinned host memory helpers (for callers without their own allocator) */
void *pms_host_alloc(uint64_t bytes);
void pms_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* PMS_H */
