"""One HSFM configuration, a few launches (for ncu / quick timing on the GPU box):
   python tools/run_case.py WORLDS PEDS STEPS [j_split [wpc [chunks [precision [reps]]]]]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from pyminisim_b200 import _capi as capi  # noqa: E402
from pyminisim_b200.batched import BatchedSimulation  # noqa: E402
from pyminisim_b200.scenarios import make_bench_crowd  # noqa: E402


def main():
    a = sys.argv[1:]
    W, N, steps = int(a[0]), int(a[1]), int(a[2])
    js = int(a[3]) if len(a) > 3 else 0
    wpc = int(a[4]) if len(a) > 4 else 0
    chunks = int(a[5]) if len(a) > 5 else 0
    precision = a[6] if len(a) > 6 else "fp32"
    reps = int(a[7]) if len(a) > 7 else 3
    sc = make_bench_crowd(W, N)
    sim = BatchedSimulation(W, N, precision=precision)
    sim.set_state(sc.poses, sc.vels)
    sim.set_speeds(sc.vmag)
    sim.set_waypoints_fixed(sc.table, loop=True)
    sim.set_robot(capi.ROBOT_UNICYCLE, sc.robot_pose, control=sc.robot_control)
    sim.set_detector(5.0, np.deg2rad(120.0), capi.DET_POLAR)
    sim.set_tuning(js, wpc, chunks, 0)
    sim.snapshot_save()
    stream = torch.cuda.ExternalStream(sim.stream)
    for rep in range(reps):
        sim.snapshot_restore()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream); sim.step_async(0.01, steps); e1.record(stream); e1.synchronize()
        print(f"{W} x {N} x {steps} j_split {js}: {e0.elapsed_time(e1):.3f} ms kernel {sim.launch_info()['kernel']}", flush=True)
    sim.close()


if __name__ == "__main__":
    main()
