"""Device-resident throughput of every BASELINE.json configuration (run on the GPU box):
  C2  1024 worlds x 32 pedestrians HSFM, 1000 steps
  C3  1 world x 65536 pedestrians HSFM (large-crowd kernel), 20 steps
  C4  4096 worlds x 64 pedestrians HSFM + detector, 1000 steps (bench.py's workload)
  C5  2048 worlds x 64 agents ORCA, 200 steps
CUDA events on the handle's stream, best of 3 after one warm-up.  Prints one JSON line per case."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from pyminisim_b200 import _capi as capi  # noqa: E402
from pyminisim_b200.batched import BatchedOrca, BatchedSimulation  # noqa: E402
from pyminisim_b200.scenarios import make_bench_crowd, make_crowd  # noqa: E402


def time_hsfm(name, W, N, steps, precision="fp32", tuning=None, large=False, reps=3, shuffle=False):
    if large:
        sc = make_crowd(W, N, side=2.0 * np.sqrt(N), local_goals=5.0)
    else:
        sc = make_bench_crowd(W, N)
    if shuffle:
        perm = np.random.default_rng(0).permutation(N)
        sc.poses, sc.table, sc.vmag, sc.vels = sc.poses[:, perm], sc.table[:, perm], sc.vmag[:, perm], sc.vels[:, perm]
    sim = BatchedSimulation(W, N, precision=precision)
    sim.set_state(sc.poses, sc.vels)
    sim.set_speeds(sc.vmag)
    sim.set_waypoints_fixed(sc.table, loop=True)
    sim.set_robot(capi.ROBOT_UNICYCLE, sc.robot_pose, control=sc.robot_control)
    sim.set_detector(5.0, np.deg2rad(120.0), capi.DET_POLAR)
    if tuning:
        sim.set_tuning(*tuning)
    sim.snapshot_save()
    stream = torch.cuda.ExternalStream(sim.stream)
    best = float("inf")
    for rep in range(reps + 1):
        sim.snapshot_restore()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream); sim.step_async(0.01, steps); e1.record(stream); e1.synchronize()
        if rep:
            best = min(best, e0.elapsed_time(e1))
    info = sim.launch_info()
    _, _, nf = sim.get_stats()
    out = dict(case=name, worlds=W, peds=N, steps=steps, precision=precision, ms=best,
               ped_updates_per_s=W * N * steps / (best * 1e-3), us_per_step=1e3 * best / steps,
               algorithmic_tflops=W * N * steps * (30.0 * N + 120.0) / (best * 1e-3) / 1e12,
               kernel=info["kernel"], blocks=info["blocks"], threads=info["threads_per_block"],
               worlds_per_block=info["worlds_per_block"], n_chunks=info["n_chunks"], nonfinite_worlds=int((nf != 0).sum()))
    sim.close()
    print(json.dumps(out), flush=True)
    return out


def time_orca(name, W, N, steps, reps=3):
    sc = make_crowd(W, N, side=16.0)
    sim = BatchedOrca(0.01, W, N)
    best = float("inf")
    for rep in range(reps + 1):
        sim.set_agents(sc.poses[..., :2], max_speeds=np.full((W, N), 1.7), pref_speeds=np.full((W, N), 1.5))
        sim.set_waypoints_fixed(sc.table, loop=True)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); sim.step(steps); e1.record(); e1.synchronize()   # pms_orca_step synchronises its own stream
        if rep:
            best = min(best, e0.elapsed_time(e1))
    out = dict(case=name, worlds=W, agents=N, steps=steps, ms=best, agent_updates_per_s=W * N * steps / (best * 1e-3))
    sim.close()
    print(json.dumps(out), flush=True)
    return out


def main():
    which = set(sys.argv[1:]) or {"c2", "c3", "c4", "c5"}
    if "c2" in which:
        time_hsfm("C2 warp kernel", 1024, 32, 1000)
        time_hsfm("C2 warp kernel, plain fp32 state", 1024, 32, 1000, precision="fp32_plain")
        time_hsfm("C2 directed-pair kernel", 1024, 32, 1000, tuning=(-1, 0, 0, 0))
        time_hsfm("C2 mixed", 1024, 32, 1000, precision="mixed")
    if "c4" in which:
        time_hsfm("C4 warp kernel", 4096, 64, 1000)
        time_hsfm("C4 warp kernel, plain fp32 state", 4096, 64, 1000, precision="fp32_plain")
        time_hsfm("C4 directed-pair kernel", 4096, 64, 1000, tuning=(-1, 0, 0, 0))
        time_hsfm("C4 mixed", 4096, 64, 1000, precision="mixed")
        time_hsfm("C4 fp64", 4096, 64, 200, precision="fp64")
        time_hsfm("4096 x 128 warp kernel", 4096, 128, 300)
        time_hsfm("4096 x 128 directed-pair kernel", 4096, 128, 300, tuning=(-1, 0, 0, 0))
        time_hsfm("16384 x 32 warp kernel", 16384, 32, 1000)
    if "c3" in which:
        time_hsfm("C3 large crowd", 1, 65536, 20, large=True)
        time_hsfm("C3 large crowd, 200 steps per call", 1, 65536, 200, large=True)
        time_hsfm("C3 large crowd, caller's order kept (no sorted working copy)", 1, 65536, 20, large=True, tuning=(0, 0, -2, 0))
        time_hsfm("C3 large crowd, every j-tile visited (no exact pruning)", 1, 65536, 20, large=True, tuning=(0, 0, -1, 0))
        time_hsfm("C3 large crowd, shuffled pedestrian order", 1, 65536, 20, large=True, shuffle=True)
        time_hsfm("C3 large crowd, shuffled pedestrian order, caller's order kept", 1, 65536, 20, large=True, shuffle=True, tuning=(0, 0, -2, 0))
        time_hsfm("C3 large crowd, MIXED", 1, 65536, 20, large=True, precision="mixed")
    if "c5" in which:
        time_orca("C5 ORCA", 2048, 64, 200)


if __name__ == "__main__":
    main()
