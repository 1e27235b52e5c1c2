// replay_api.cuh -- host-side entry points of the recorded-trajectory playback kernel (replay.cu).
#pragma once
#include "pms_common.cuh"

namespace pms {

// ReplayPedestriansPolicy (pedestrians/_replay.py:19-59): frames (T, W, N, 3) of poses resident in HBM
struct ReplayArgs {
    const double *poses;    // (n_frames, W, N, 3)
    int n_frames, loop;
    int frame0;             // frame shown before the first step of the launch
};

// n_steps of: robot step (device model / external pose) -> next frame (_replay.py:46-52) -> collision list + detector on
// the frame's poses (+ the sensor plan's records).  One warp per world; sens = a.sp is in use.
cudaError_t replay_launch(const HsfmArgs &a, const ReplayArgs &ra, bool sens, cudaStream_t stream);

// frame shown after n_steps more steps (the rule of _replay.py:46-52)
inline int replay_advance(int frame, int n_frames, int loop, long long n_steps)
{
    const int last = n_frames - 1;
    if (n_steps <= 0 || n_frames <= 1) return frame;
    if (!loop) return (long long)frame + n_steps >= last ? last : frame + (int)n_steps;
    return (int)(((long long)frame + n_steps) % n_frames);
}

} // namespace pms
