"""Per-step latency of the drop-in `Simulation.step()` (SURVEY.md section 6 measured the reference the same way):
HSFM + RandomWaypointTracker + unicycle robot + default PedestrianDetector on an EmptyWorld, rt_factor=None, no renderer,
N = 2 (examples/basic.py) / 8 / 32 / 64 pedestrians, one world, one step per call.

    python tools/bench_dropin.py                 # this package (needs a GPU)
    python tools/bench_dropin.py --reference     # the reference itself (build container only: imports /root/reference)

Prints one JSON line per N.  Positions are drawn far enough apart for the reference's Euler scheme to stay finite."""
import json
import sys
import time
import types

import numpy as np

REFERENCE = "--reference" in sys.argv
if REFERENCE:
    sys.modules.setdefault("rvo2", types.ModuleType("rvo2"))       # pyminisim.pedestrians imports it at module level
    sys.path.insert(0, "/root/reference")
    from pyminisim.core import Simulation
    from pyminisim.pedestrians import HeadedSocialForceModelPolicy, RandomWaypointTracker
    from pyminisim.robot import UnicycleRobotModel
    from pyminisim.sensors import PedestrianDetector
    from pyminisim.world_map import EmptyWorld
else:
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from pyminisim_b200.core import Simulation
    from pyminisim_b200.pedestrians import HeadedSocialForceModelPolicy, RandomWaypointTracker
    from pyminisim_b200.robot import UnicycleRobotModel
    from pyminisim_b200.sensors import PedestrianDetector
    from pyminisim_b200.world_map import EmptyWorld


def run(n, steps, warm):
    Simulation.seed(7)
    side = max(7.0, 5.0 * np.sqrt(n))
    g = int(np.ceil(np.sqrt(n)))
    xs = (np.arange(g) + 0.5) / g * side - side / 2
    pos = np.array([[x, y] for x in xs for y in xs])[:n]
    poses = np.concatenate([pos, np.zeros((n, 1))], axis=1)
    tracker = RandomWaypointTracker(world_size=(side, side))
    model = HeadedSocialForceModelPolicy(n_pedestrians=n, waypoint_tracker=tracker, initial_poses=poses)
    sim = Simulation(world_map=EmptyWorld(),
                     robot_model=UnicycleRobotModel(initial_pose=np.array([-side / 2 - 1.0, 0.0, 0.0]),
                                                    initial_control=np.array([0.5, 0.0])),
                     pedestrians_model=model, sensors=[PedestrianDetector()], sim_dt=0.01, rt_factor=None)
    for _ in range(warm):
        sim.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        st = sim.step()
    dt = time.perf_counter() - t0
    p = st.world.pedestrians.poses[0]
    return dict(impl="reference" if REFERENCE else "pyminisim_b200", n_pedestrians=n, steps=steps,
                us_per_step=1e6 * dt / steps, ped_updates_per_s=n * steps / dt, finite=bool(np.isfinite(p).all()))


def run_orca(n, steps, warm):
    """ORCAPedestriansModel in the same loop (this package only: the reference's ORCA needs the absent rvo2).  The sensor
    pass is the stand-alone one (pms_sense_tick), as for every model without a fused epilogue."""
    from pyminisim_b200.pedestrians import FixedWaypointTracker, ORCAPedestriansModel
    side = max(7.0, 2.0 * np.sqrt(n))
    g = int(np.ceil(np.sqrt(n)))
    xs = (np.arange(g) + 0.5) / g * side - side / 2
    pos = np.array([[x, y] for x in xs for y in xs])[:n]
    tracker = FixedWaypointTracker(initial_positions=pos.copy(), waypoints=np.stack([-pos, pos], axis=1), loop=True)
    model = ORCAPedestriansModel(dt=0.01, waypoint_tracker=tracker, n_pedestrians=n,
                                 initial_poses=np.concatenate([pos, np.zeros((n, 1))], axis=1))
    sim = Simulation(world_map=EmptyWorld(),
                     robot_model=UnicycleRobotModel(initial_pose=np.array([-side / 2 - 1.0, 0.0, 0.0]),
                                                    initial_control=np.array([0.5, 0.0])),
                     pedestrians_model=model, sensors=[PedestrianDetector()], sim_dt=0.01, rt_factor=None)
    for _ in range(warm):
        sim.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.step()
    dt = time.perf_counter() - t0
    return dict(impl="pyminisim_b200", model="orca", n_pedestrians=n, steps=steps, us_per_step=1e6 * dt / steps)


if __name__ == "__main__":
    if "--profile" in sys.argv:          # where the Python side of a tick goes (cProfile, N = 64); `--profile orca`: the ORCA model
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        (run_orca if "orca" in sys.argv else run)(64, 2000, 50)
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(22)
        sys.exit(0)
    for n in (2, 8, 32, 64):
        print(json.dumps(run(n, 300 if REFERENCE and n >= 32 else 1000, 50)), flush=True)
    if not REFERENCE:
        for n in (8, 64):
            print(json.dumps(run_orca(n, 1000, 50)), flush=True)
