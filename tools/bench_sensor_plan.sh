# C4 (4096 worlds x 64 pedestrians, 1000 fused steps) with a CirclesWorld of 8 obstacles per world: cost of the sensor plans
python - <<'PY'
# cost of the reference's sensors inside the C4 launch
import numpy as np, torch, sys
sys.path.insert(0, '.')
from pyminisim_b200 import _capi as capi
from pyminisim_b200.batched import BatchedSimulation
from pyminisim_b200.scenarios import make_bench_crowd
W, N, K = 4096, 64, 1000
sc = make_bench_crowd(W, N)
rng = np.random.default_rng(5)
circles = np.concatenate([rng.uniform(-25, 25, (W, 8, 2)), rng.uniform(0.3, 1.5, (W, 8, 1))], axis=2)
for plan in ("none", "lidar 10 Hz", "omni every step", "detector masks + collision list every step", "lidar + omni + detector masks + collision list every step", "all + detector values every 11 steps"):
    sim = BatchedSimulation(W, N, precision="fp32")
    sim.set_state(sc.poses, sc.vels); sim.set_speeds(sc.vmag); sim.set_waypoints_fixed(sc.table, loop=True)
    sim.set_robot(capi.ROBOT_UNICYCLE, sc.robot_pose, control=sc.robot_control)
    sim.set_detector(5.0, np.deg2rad(120.0), capi.DET_POLAR); sim.set_map(capi.MAP_CIRCLES, circles)
    if plan.startswith("lidar") or plan.startswith("all"):
        sim.set_lidar(100, 8.0, 0.01); sim.schedule_sensor(capi.SENSOR_LIDAR, 0.1)
    if plan.startswith("omni") or plan.startswith("lidar +") or plan.startswith("all"):
        sim.schedule_sensor(capi.SENSOR_OMNI, 0.0)
    if plan.startswith("detector") or plan.startswith("lidar +") or plan.startswith("all"):
        sim.schedule_sensor(capi.SENSOR_COLLISIONS, 0.0)
        sim.schedule_sensor(capi.SENSOR_DETECTOR, 0.1 if plan.startswith("all") else 0.0, values=plan.startswith("all"))
    sim.snapshot_save()
    st = torch.cuda.ExternalStream(sim.stream)
    best = 1e9
    for rep in range(4):
        sim.snapshot_restore()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st); sim.step_async(0.01, K); e1.record(st); e1.synchronize()
        if rep: best = min(best, e0.elapsed_time(e1))
    print(f"C4 with sensor plan [{plan}]: {best:.3f} ms per 1000 steps, kernel {sim.launch_info()['kernel']}", flush=True)
    sim.close()
PY
