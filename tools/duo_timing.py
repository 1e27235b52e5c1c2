"""Per-role busy / barrier-wait cycles per step of the pair-warp / owner-warp kernel (experiment build with -DPMS_DUO_TIMING):
   PMS_LIB=pyminisim_b200/libpms_timing.so python tools/duo_timing.py WORLDS PEDS [STEPS]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyminisim_b200 import _capi as capi  # noqa: E402
from pyminisim_b200.batched import BatchedSimulation  # noqa: E402
from pyminisim_b200.scenarios import make_bench_crowd  # noqa: E402

W, N = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
sc = make_bench_crowd(W, N)
sim = BatchedSimulation(W, N, precision="fp32")
sim.set_state(sc.poses, sc.vels); sim.set_speeds(sc.vmag); sim.set_waypoints_fixed(sc.table, loop=True)
sim.set_robot(capi.ROBOT_UNICYCLE, sc.robot_pose, control=sc.robot_control)
sim.set_detector(5.0, np.deg2rad(120.0), capi.DET_POLAR)
sim.set_tuning(4)
sim.snapshot_save()
sim.step(0.01, steps)
sim.pop_waypoint_events()
sim.snapshot_restore()
sim.step(0.01, steps)
ev = sim.pop_waypoint_events()
ev = np.asarray(ev)
for wid in sorted(set(ev[:, 2])):
    m = ev[ev[:, 2] == wid]
    print(f"{W} x {N} warp {wid}: busy {m[:, 0].mean():7.0f} (max {m[:, 0].max()}) wait {m[:, 1].mean():7.0f} cycles per step")
