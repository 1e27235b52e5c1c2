// orca.cuh -- interface of the ORCA velocity-obstacle solve (implemented in orca.cu, which is compiled
// with -fmad=false so that its fp32 arithmetic rounds exactly like the RVO2 library's scalar code).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pms {

constexpr int ORCA_MAX_NEIGHBORS = 16;

struct OrcaState {
    int W = 0, N = 0, Np = 0;
    // ORCAParams (pedestrians/_orca.py:13-21) / PyRVOSimulator ctor (:73-79)
    float time_step = 0.01f, neighbor_dist = 1.f, time_horizon = 2.f, radius = 0.3f;
    int max_neighbors = 10;
    double goal_threshold = 0.2;
    // rvo2 agent state (single precision like the library)
    float2 *pos = nullptr, *vel = nullptr, *pref = nullptr;
    float *max_speed = nullptr;
    double *pref_speed = nullptr;
    // FixedWaypointTracker
    double2 *cur_wp = nullptr, *table = nullptr;
    int *wp_index = nullptr;
    int n_table = 0, loop = 0;
    double reach = 0.3;
    double *heading = nullptr;
    int *n_constraints = nullptr;
    double *stage = nullptr;
    size_t stage_elems = 0;
    bool agents_set = false, waypoints_set = false;
    bool refresh_pref = true;   // waypoint setters recompute the preferred velocities (off: ORCAPedestriansModel.reset_to_state)
};

int orca_alloc(OrcaState &s, int W, int N, double time_step, double neighbor_dist, int max_neighbors,
               double time_horizon, double radius, double goal_threshold);
void orca_free(OrcaState &s);
int orca_set_agents(OrcaState &s, const double *pos, const double *vel, const double *max_speeds,
                    const double *pref_speeds, cudaStream_t stream);
int orca_set_waypoints(OrcaState &s, const double *table, int n_rows, const int32_t *start_index, double reach,
                       int loop, cudaStream_t stream);
int orca_set_current_waypoints(OrcaState &s, const double *current, cudaStream_t stream);
int orca_step(OrcaState &s, int n_steps, cudaStream_t stream, long long *launches);
int orca_get_agents(OrcaState &s, double *poses, double *vels, int32_t *n_constraints, cudaStream_t stream);
int orca_get_agents_async(OrcaState &s, double *pinned /*(6 W N) host-pinned*/, cudaStream_t stream);

} // namespace pms
