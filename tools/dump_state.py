"""Final state of a few scenes with the library PMS_LIB points at -> npz (A/B of two builds: tools/dump_state.py out.npz)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyminisim_b200 import _capi as capi  # noqa: E402
from pyminisim_b200.batched import BatchedSimulation  # noqa: E402
from pyminisim_b200.scenarios import make_bench_crowd, make_crowd  # noqa: E402

out = {}
for name, sc, prec, loop in (("bench64", make_bench_crowd(96, 64), "fp32", True), ("bench64p", make_bench_crowd(96, 64), "fp32_plain", True),
                             ("mirror64", make_crowd(96, 64), "fp32", True), ("mirror50", make_crowd(64, 50), "fp32", False),
                             ("bench128", make_bench_crowd(32, 128), "fp32", True), ("bench32", make_bench_crowd(64, 32), "fp32", True),
                             ("mirror20", make_crowd(64, 20), "fp32", True)):
    if len(sys.argv) > 2 and int(sys.argv[2]) == 4 and sc.poses.shape[1] > 64:
        continue
    W, N = sc.poses.shape[:2]
    sim = BatchedSimulation(W, N, precision=prec)
    sim.set_state(sc.poses, sc.vels); sim.set_speeds(sc.vmag); sim.set_waypoints_fixed(sc.table, loop=loop)
    sim.set_robot(capi.ROBOT_UNICYCLE, sc.robot_pose, control=sc.robot_control)
    sim.set_detector(5.0, np.deg2rad(120.0), capi.DET_POLAR)
    sim.set_tuning(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
    sim.step(0.01, 700)
    p, v = sim.get_state()
    m, vals = sim.get_detector()
    det, col, nf = sim.get_stats()
    out.update({name + "_p": p, name + "_v": v, name + "_m": m, name + "_vals": vals, name + "_det": det, name + "_col": col, name + "_nf": nf,
                name + "_wp": sim.get_waypoints()[1]})
    sim.close()
np.savez(sys.argv[1], **out)
