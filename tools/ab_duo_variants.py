"""ms per 1000 steps of the pair-warp / owner-warp kernel's variants (and of the automatic choice) on W x 64 batches:
   python tools/ab_duo_variants.py 128 256 512 ..."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyminisim_b200 import _capi as capi  # noqa: E402
from pyminisim_b200.batched import BatchedSimulation  # noqa: E402
from pyminisim_b200.scenarios import make_bench_crowd  # noqa: E402

N = 64
for W in [int(x) for x in sys.argv[1:]]:
    sc = make_bench_crowd(W, N)
    row = []
    for tune in ((4, 1), (4, 2), (4, 3), (4, 4), (4, 5), (5, 0), (0, 0)):      # the kernel's five variants, warp-per-world kernel, automatic
        sim = BatchedSimulation(W, N, precision="fp32")
        sim.set_state(sc.poses, sc.vels); sim.set_speeds(sc.vmag); sim.set_waypoints_fixed(sc.table, loop=True)
        sim.set_robot(capi.ROBOT_UNICYCLE, sc.robot_pose, control=sc.robot_control)
        sim.set_detector(5.0, np.deg2rad(120.0), capi.DET_POLAR)
        sim.set_tuning(*tune)
        sim.snapshot_save()
        stream = torch.cuda.ExternalStream(sim.stream)
        ts = []
        for _ in range(6):
            sim.snapshot_restore()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); sim.step_async(0.01, 1000); e1.record(stream); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        info = sim.launch_info()
        row.append(f"{tune}: {np.median(ts[2:]):.3f} (k{info['kernel']} v{info['j_split']} occ{info['blocks_per_sm']})")
        sim.close()
    print(f"{W} x {N}: " + "  ".join(row), flush=True)
