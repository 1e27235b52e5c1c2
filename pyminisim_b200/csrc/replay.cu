// replay.cu -- ReplayPedestriansPolicy (pedestrians/_replay.py) as a batched gather on the device: the recorded frames stay
// in HBM, a fused launch plays n_steps of them under the robot and the sensors.  One warp per world: the robot a block of
// 32 steps ahead (warp_robot_block of hsfm_warp.cuh, which also takes the sensor plan's lidar / omni readings), then per step
// the frame's poses against the collision list (core/_simulation.py:141-143) and the pedestrian detector
// (sensors/_pedestrian_detector.py:86-134) with the exact fp64 predicates.
#include "hsfm_warp.cuh"
#include "replay_api.cuh"

namespace pms {
namespace {

constexpr int REPLAY_WPC = 4;           // world warps per CTA

template <bool SENS>
__global__ void __launch_bounds__(32 * REPLAY_WPC) replay_kernel(const __grid_constant__ HsfmArgs a, const ReplayArgs ra)
{
    __shared__ WarpRobotSlot s_ring[REPLAY_WPC][WARP_ROBOT_BLOCK];
    __shared__ RobotRegs s_regs[REPLAY_WPC];
    const int lane = threadIdx.x & 31;
    const int wid = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int w = blockIdx.x * REPLAY_WPC + wid;
    if (w >= a.W) return;
    const int N = a.N, words = (N + 31) >> 5;
    const bool has_robot = a.robot_model != ROBOT_NONE;
    WarpRobotSlot *ring = s_ring[wid];
    RobotRegs *rrp = &s_regs[wid];
    if (has_robot && lane == 0) robot_load(a, w, *rrp);
    __syncwarp();
    unsigned det_acc = 0, coll_acc = 0;
    int f = ra.frame0;
    const int f_last = ra.n_frames - 1;
    for (int t = 0; t < a.n_steps; ++t) {
        if (has_robot && (t & (WARP_ROBOT_BLOCK - 1)) == 0)
            warp_robot_block<SENS>(a, w, t, min(WARP_ROBOT_BLOCK, a.n_steps - t), rrp, ring, lane);
        f = f == f_last ? (ra.loop ? 0 : f) : f + 1;                    // _replay.py:46-52
        if (!has_robot) continue;
        const WarpRobotSlot &rb = ring[t & (WARP_ROBOT_BLOCK - 1)];
        const RobotSlot r64{rb.x, rb.y, rb.th, 0.0, 0.0};
        const bool last = t == a.n_steps - 1;
        int dslot = -1, cslot = -1;
        if constexpr (SENS) {
            if (a.sp.ped_on) {
                if (a.sp.fire[SENSOR_DETECTOR]) dslot = a.sp.fire[SENSOR_DETECTOR][t];
                if (a.sp.fire[SENSOR_COLLISIONS]) cslot = a.sp.fire[SENSOR_COLLISIONS][t];
            }
        }
        const bool rec_vals = SENS && dslot >= 0 && a.sp.det_vals != nullptr;
        const double *frame = ra.poses + ((size_t)f * a.W + w) * N * 3;
        for (int i0 = 0; i0 < N; i0 += 32) {
            const int i = i0 + lane;
            SenseOut so{false, false, 0.0, 0.0};
            if (i < N) so = sense_ped<double>(a, r64, frame[3 * i], frame[3 * i + 1], last || rec_vals);
            const unsigned dm = __ballot_sync(0xffffffffu, so.det);
            const unsigned cm = __ballot_sync(0xffffffffu, so.coll);
            det_acc += __popc(dm);
            coll_acc += __popc(cm);
            if (i < N) {
                if (last && a.det_enabled) a.det_vals[(size_t)w * N + i] = make_double2(so.v0, so.v1);
                if constexpr (SENS) {
                    if (rec_vals) a.sp.det_vals[((size_t)dslot * a.W + w) * N + i] = make_double2(so.v0, so.v1);
                }
            }
            if (lane == 0) {
                if (last) {
                    a.det_mask[(size_t)w * words + (i0 >> 5)] = dm;
                    a.coll_mask[(size_t)w * words + (i0 >> 5)] = cm;
                }
                if constexpr (SENS) {
                    if (dslot >= 0) a.sp.det_mask[((size_t)dslot * a.W + w) * words + (i0 >> 5)] = dm;
                    if (cslot >= 0) a.sp.coll_mask[((size_t)cslot * a.W + w) * words + (i0 >> 5)] = cm;
                }
            }
        }
    }
    __syncwarp();
    if (has_robot && lane == 0) {
        robot_store(a, w, *rrp);
        if (det_acc) atomicAdd(&a.det_count[w], (unsigned long long)det_acc);
        if (coll_acc) atomicAdd(&a.coll_count[w], (unsigned long long)coll_acc);
    }
}

} // namespace

cudaError_t replay_launch(const HsfmArgs &a, const ReplayArgs &ra, bool sens, cudaStream_t stream)
{
    const int blocks = (a.W + REPLAY_WPC - 1) / REPLAY_WPC;
    if (sens) replay_kernel<true><<<blocks, 32 * REPLAY_WPC, 0, stream>>>(a, ra);
    else replay_kernel<false><<<blocks, 32 * REPLAY_WPC, 0, stream>>>(a, ra);
    return cudaGetLastError();
}

} // namespace pms
