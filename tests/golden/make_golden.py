"""Generate golden fixtures from the LIVE reference (TimeEscaper/pyminisim at /root/reference).

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference is imported unmodified; the only shim is an empty ``rvo2`` module so that
``pyminisim.pedestrians`` (which imports ``_orca`` -> ``rvo2``) can be imported at all.  Outputs are
small ``.npz`` files committed beside this script; tests compare the CPU oracle and the CUDA path
against them.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.modules.setdefault("rvo2", types.ModuleType("rvo2"))

from pyminisim.core import Simulation  # noqa: E402
from pyminisim.pedestrians import (FixedWaypointTracker, HeadedSocialForceModelPolicy,  # noqa: E402
                                   RandomWaypointTracker)
from pyminisim.pedestrians._hsfm import _calc_f_p_ij  # noqa: E402
from pyminisim.robot import BicycleRobotModel, UnicycleRobotModel  # noqa: E402
from pyminisim.sensors import (LidarSensor, LidarSensorConfig, OmniObstacleDetector,  # noqa: E402
                               PedestrianDetector, PedestrianDetectorConfig)
from pyminisim.util import wrap_angle  # noqa: E402
from pyminisim.world_map import AABBWorld, CirclesWorld, EmptyWorld  # noqa: E402
from pyminisim.world_map._aabb_world import AABBObject  # noqa: E402

from pyminisim_b200.scenarios import make_crowd  # noqa: E402

RETURN_TYPES = ["polar", "relative_rotated", "relative", "absolute"]


def _mask(ids, n):
    m = np.zeros((n + 31) // 32, dtype=np.uint32)
    for i in ids:
        m[i >> 5] |= np.uint32(1 << (i & 31))
    return m


def _det_arrays(reading, n):
    vals = np.zeros((n, 2))
    for k, v in reading.pedestrians.items():
        vals[k] = v
    return _mask(reading.pedestrians.keys(), n), vals


def _ped_arrays(state):
    n = len(state.poses)
    poses = np.stack([state.poses[i] for i in range(n)])
    vels = np.stack([state.velocities[i] for i in range(n)])
    return poses, vels


def gen_pair_forces():
    rng = np.random.default_rng(7)
    n = 256
    r_i = rng.uniform(-3, 3, (n, 2))
    r_j = r_i + rng.uniform(-1.5, 1.5, (n, 2))
    r_j[:8] = r_i[:8] + rng.uniform(-0.3, 0.3, (8, 2))  # overlapping pairs
    v_i = rng.uniform(-2, 2, (n, 2))
    v_j = rng.uniform(-2, 2, (n, 2))
    rad_j = np.where(np.arange(n) % 2 == 0, 0.3, 0.35)
    out = np.stack([_calc_f_p_ij(0.3, rad_j[k], r_i[k], r_j[k], v_i[k], v_j[k], 2000, 0.08, 1.2e5, 2.4e5)
                    for k in range(n)])
    kat1 = np.stack([
        _calc_f_p_ij(0.3, 0.3, np.array([0., 0.]), np.array([0.5, 0.]), np.array([1., 0.]), np.array([0., 1.]),
                     2000, 0.08, 1.2e5, 2.4e5),
        _calc_f_p_ij(0.3, 0.3, np.array([0., 0.]), np.array([0.6, 0.8]), np.array([1., 0.]), np.array([0., 1.]),
                     2000, 0.08, 1.2e5, 2.4e5)])
    angles = rng.uniform(-20, 20, 64)
    wrapped = np.array([wrap_angle(a) for a in angles])
    np.savez(os.path.join(HERE, "pair_forces.npz"), r_i=r_i, r_j=r_j, v_i=v_i, v_j=v_j, rad_j=rad_j, f=out,
             kat1=kat1, angles=angles, wrapped=wrapped)


def run_sim(poses, table, robot_pose, robot_control, horizons, vmag=1.5, loop=True, det=(5.0, 120.0, "polar"),
            robot="unicycle", robot_visible=True, world_map=None, controls=None, vels=None, wheel_base=0.324):
    """One reference Simulation; records state at the given step counts."""
    n = poses.shape[0]
    tracker = FixedWaypointTracker(initial_positions=table[:, 0, :].copy(), waypoints=table[:, 1:, :].copy(), loop=loop)
    model = HeadedSocialForceModelPolicy(waypoint_tracker=tracker, n_pedestrians=n, initial_poses=poses.copy(),
                                         initial_velocities=None if vels is None else vels.copy(),
                                         pedestrian_linear_velocity_magnitude=vmag, robot_visible=robot_visible)
    if robot == "unicycle":
        rm = UnicycleRobotModel(initial_pose=robot_pose.copy(), initial_control=robot_control.copy())
    elif robot == "bicycle":
        rm = BicycleRobotModel(wheel_base=wheel_base, initial_center_pose=robot_pose.copy(),
                               initial_control=robot_control.copy())
    else:
        rm = None
    sensors = [PedestrianDetector(config=PedestrianDetectorConfig(max_dist=det[0], fov=np.deg2rad(det[1]),
                                                                  return_type=det[2]))]
    sim = Simulation(world_map=world_map or EmptyWorld(), robot_model=rm, pedestrians_model=model, sensors=sensors,
                     sim_dt=0.01, rt_factor=None)
    rec = {k: [] for k in ("poses", "vels", "robot_pose", "robot_vel", "det_mask", "det_vals", "coll_mask",
                           "wp_index", "robot_world_collision")}
    det_count, coll_count = 0, 0
    det_counts, coll_counts = [], []
    for step in range(1, max(horizons) + 1):
        ctrl = None if controls is None else controls[min(step - 1, len(controls) - 1)]
        st = sim.step(ctrl)
        if rm is not None:
            det_count += len(st.sensors["pedestrian_detector"].reading.pedestrians)
            coll_count += len(st.world.robot_to_pedestrians_collisions)
        if step in horizons:
            p, v = _ped_arrays(st.world.pedestrians)
            rec["poses"].append(p)
            rec["vels"].append(v)
            if rm is not None:
                rec["robot_pose"].append(st.world.robot.pose)
                rec["robot_vel"].append(st.world.robot.velocity)
                m, vals = _det_arrays(st.sensors["pedestrian_detector"].reading, n)
                rec["det_mask"].append(m)
                rec["det_vals"].append(vals)
                rec["coll_mask"].append(_mask(st.world.robot_to_pedestrians_collisions, n))
            rec["robot_world_collision"].append(int(st.world.robot_to_world_collision))
            rec["wp_index"].append(np.array([tracker.state.current_indices[i] for i in range(n)], dtype=np.int32))
            det_counts.append(det_count)
            coll_counts.append(coll_count)
    out = {k: np.array(v) for k, v in rec.items() if len(v)}
    out["det_count"] = np.array(det_counts)
    out["coll_count"] = np.array(coll_counts)
    out["horizons"] = np.array(horizons)
    return out


def gen_kat2():
    poses = np.array([[2., .5, 0.], [2., -.5, np.pi], [-1., 2., 1.]])
    wps = np.array([[[-3., .5], [3., .5]], [[-3., -.5], [3., -.5]], [[-1., -3.], [-1., 3.]]])
    table = np.concatenate([poses[:, None, :2], wps], axis=1)
    out = run_sim(poses, table, np.array([0., 0., 0.]), np.array([0.5, 0.2]), [1, 100, 300])
    np.savez(os.path.join(HERE, "kat2.npz"), init_poses=poses, table=table, **out)


def gen_crowds():
    for n_peds, n_worlds, horizons in ((8, 3, [1, 100, 300, 1000]), (32, 2, [1, 100, 300, 1000]),
                                       (64, 2, [1, 100, 300, 1000])):
        sc = make_crowd(n_worlds, n_peds)
        outs = [run_sim(sc.poses[w], sc.table[w], sc.robot_pose[w], sc.robot_control[w], horizons)
                for w in range(n_worlds)]
        merged = {k: np.stack([o[k] for o in outs]) for k in outs[0] if k != "horizons"}
        np.savez(os.path.join(HERE, f"crowd_n{n_peds}.npz"), horizons=np.array(horizons), side=sc.side, **merged)
        print("crowd", n_peds, "finite:", np.isfinite(merged["poses"]).all())


def gen_variants():
    """Return types, invisible robot, non-loop freeze, bicycle, external controls, per-ped speeds."""
    sc = make_crowd(1, 8, side=5.0, first_world=50)
    res = {}
    for rt in RETURN_TYPES:
        o = run_sim(sc.poses[0], sc.table[0], np.array([-1.0, 0.2, 0.3]), np.array([0.4, 0.3]), [1, 50, 150],
                    det=(4.0, 200.0, rt))
        res[f"rt_{rt}_det_mask"] = o["det_mask"]
        res[f"rt_{rt}_det_vals"] = o["det_vals"]
        res[f"rt_{rt}_poses"] = o["poses"]
    o = run_sim(sc.poses[0], sc.table[0], np.array([-1.0, 0.2, 0.3]), np.array([0.4, 0.3]), [1, 50, 150],
                robot_visible=False)
    res["invisible_poses"], res["invisible_vels"] = o["poses"], o["vels"]
    o = run_sim(sc.poses[0], sc.table[0], None, None, [1, 50, 150], robot=None)
    res["norobot_poses"], res["norobot_vels"] = o["poses"], o["vels"]
    # non-loop: pedestrians freeze at their last waypoint (_hsfm.py:302-304)
    vm = np.linspace(1.0, 2.2, 8)
    o = run_sim(sc.poses[0], sc.table[0, :, :2], np.array([-4.0, 0.0, 0.0]), np.array([0.3, 0.0]),
                [1, 200, 400, 700], loop=False, vmag=vm)
    for k in ("poses", "vels", "wp_index", "det_mask", "coll_mask"):
        res[f"freeze_{k}"] = o[k]
    res["freeze_vmag"] = vm
    # bicycle robot (velocity reported as zero) with a changing control sequence
    ctr = np.stack([np.array([0.8, 0.3 * np.sin(0.05 * k)]) for k in range(150)])
    o = run_sim(sc.poses[0], sc.table[0], np.array([-1.0, 0.2, 0.3]), np.array([0.0, 0.0]), [1, 50, 150],
                robot="bicycle", controls=ctr)
    for k in ("poses", "vels", "robot_pose", "robot_vel", "det_mask"):
        res[f"bicycle_{k}"] = o[k]
    res["bicycle_controls"] = ctr
    o = run_sim(sc.poses[0], sc.table[0], np.array([-1.0, 0.2, 0.3]), np.array([0.0, 0.0]), [1, 50, 150],
                robot="unicycle", controls=ctr)
    for k in ("poses", "vels", "robot_pose", "robot_vel", "det_mask", "coll_mask"):
        res[f"unictrl_{k}"] = o[k]
    # robot blocked by a circle obstacle (move vetoed, control not updated)
    world = CirclesWorld(circles=np.array([[1.0, 0.2, 0.5], [-3.0, 3.0, 0.7]]))
    ctr2 = np.stack([np.array([1.0, 0.1 * (k % 7)]) for k in range(200)])
    o = run_sim(sc.poses[0], sc.table[0], np.array([-1.0, 0.2, 0.0]), np.array([0.0, 0.0]), [1, 60, 120, 200],
                robot="unicycle", controls=ctr2, world_map=world)
    for k in ("poses", "robot_pose", "robot_vel", "robot_world_collision"):
        res[f"blocked_{k}"] = o[k]
    res["blocked_controls"] = ctr2
    res["blocked_circles"] = world.circles
    np.savez(os.path.join(HERE, "variants.npz"), init_poses=sc.poses[0], table=sc.table[0], **res)


def gen_sensors():
    res = {}
    circles = np.array([[2.0, 2.0, 1.0], [-3.0, 1.0, 0.5], [0.5, -2.5, 0.8], [6.0, -6.0, 2.0], [-1.0, -1.0, 0.3]])
    boxes = [(3.0, 1.0, 2.0, 1.5), (-2.0, -4.0, 1.0, 3.0), (0.5, 3.0, 0.5, 0.5), (-4.0, 2.0, 3.0, 0.2)]
    robots = np.array([[0., 0., 0.], [1.0, 1.2, 0.4], [-2.5, 0.7, -1.0], [5.0, -3.0, 2.0], [2.2, 2.1, 0.0],
                       [-7.5, 7.5, 1.0]])
    cw = CirclesWorld(circles=circles)
    aw = AABBWorld([AABBObject(box=b) for b in boxes])

    class _R:  # minimal duck-typed robot / world state for the sensors
        def __init__(self, pose):
            self.pose = pose

    class _W:
        def __init__(self, pose):
            self.robot = _R(pose)
            self.pedestrians = None

    for name, world in (("circles", cw), ("aabb", aw), ("empty", EmptyWorld())):
        for cfgname, cfg in (("default", LidarSensorConfig()), ("coarse", LidarSensorConfig(n_beams=37, max_dist=5.0, resolution=0.03))):
            lid = LidarSensor(config=cfg)
            pts_all, cnt = [], []
            for r in robots:
                pts = lid.get_reading(_W(r), world).points
                pts = pts.reshape(-1, 2)
                cnt.append(len(pts))
                pad = np.zeros((cfg.n_beams, 2))
                pad[:len(pts)] = pts
                pts_all.append(pad)
            res[f"lidar_{name}_{cfgname}_points"] = np.stack(pts_all)
            res[f"lidar_{name}_{cfgname}_count"] = np.array(cnt)
        omni = OmniObstacleDetector()
        res[f"omni_{name}"] = np.array([omni.get_reading(_W(r), world).min_obstacle_dist for r in robots])
        res[f"collision_{name}"] = np.array([bool(world.is_collision(r[:2].copy(), 0.35)) for r in robots])
    res["circles"], res["boxes"], res["robots"] = circles, np.array(boxes), robots
    # scheduler quirk: default 10 Hz lidar fires at 0-based steps 10, 21, 32
    calls = []

    class CountingLidar(LidarSensor):
        def get_reading(self, ws, wm):
            calls.append(step_box[0])
            return super().get_reading(ws, wm)

    step_box = [-1]
    sim = Simulation(world_map=cw, robot_model=UnicycleRobotModel(np.array([0., 0., 0.])), pedestrians_model=None,
                     sensors=[CountingLidar()], sim_dt=0.01, rt_factor=None)
    for k in range(40):
        step_box[0] = k
        sim.step()
    res["lidar_fire_steps"] = np.array(calls)
    np.savez(os.path.join(HERE, "sensors.npz"), **res)


def gen_random_tracker():
    """KAT-3: examples/basic.py set-up, Simulation.seed(42) before constructing anything."""
    Simulation.seed(42)
    robot_model = UnicycleRobotModel(initial_pose=np.array([0., 0., 0.0]),
                                     initial_control=np.array([0.0, np.deg2rad(25.0)]))
    tracker = RandomWaypointTracker(world_size=(7.0, 7.0))
    init = np.array([[-3., -3., 0.], [3., 3., 0.]])
    model = HeadedSocialForceModelPolicy(n_pedestrians=2, waypoint_tracker=tracker,
                                         pedestrian_linear_velocity_magnitude=np.array([1.5, 2.5]),
                                         initial_poses=init.copy())
    cur0 = np.stack([tracker.state.current_waypoints[i] for i in range(2)])
    nxt0 = np.stack([tracker.state.next_waypoints[i] for i in range(2)])
    sim = Simulation(world_map=EmptyWorld(), robot_model=robot_model, pedestrians_model=model,
                     sensors=[PedestrianDetector()], rt_factor=None)
    horizons = [1, 100, 500, 1001, 3000]
    poses, vels, curs = [], [], []
    for step in range(1, max(horizons) + 1):
        st = sim.step()
        if step in horizons:
            p, v = _ped_arrays(st.world.pedestrians)
            poses.append(p)
            vels.append(v)
            curs.append(np.stack([tracker.state.current_waypoints[i] for i in range(2)]))
    np.savez(os.path.join(HERE, "random_tracker.npz"), init_poses=init, cur0=cur0, next0=nxt0, horizons=np.array(horizons),
             poses=np.array(poses), vels=np.array(vels), cur=np.array(curs), vmag=np.array([1.5, 2.5]))


def gen_replay():
    """ReplayPedestriansPolicy (pedestrians/_replay.py) under a Simulation with a unicycle robot and the detector:
    recorded trajectories of 5 pedestrians over 12 frames, loop and no-loop, 30 simulation steps each."""
    from pyminisim.pedestrians import ReplayPedestriansPolicy
    rng = np.random.default_rng(21)
    T, n = 12, 5
    start = rng.uniform(-2.5, 2.5, (n, 2))
    vel = rng.uniform(-1.0, 1.0, (n, 2))
    poses = np.zeros((T, n, 3)); vels = np.zeros((T, n, 3))
    for t in range(T):
        poses[t, :, :2] = start + vel * (0.1 * t)
        poses[t, :, 2] = np.arctan2(vel[:, 1], vel[:, 0])
        vels[t, :, :2] = vel
    res = {"rec_poses": poses, "rec_vels": vels}
    for loop in (False, True):
        model = ReplayPedestriansPolicy(poses.copy(), vels.copy(), loop=loop)
        rm = UnicycleRobotModel(initial_pose=np.array([-1.0, 0.2, 0.3]), initial_control=np.array([0.8, 0.4]))
        sensors = [PedestrianDetector(config=PedestrianDetectorConfig(max_dist=4.0, fov=np.deg2rad(140.0), return_type="relative"))]
        sim = Simulation(world_map=EmptyWorld(), robot_model=rm, pedestrians_model=model, sensors=sensors, sim_dt=0.01, rt_factor=None)
        steps, ps, masks, vals, colls = [], [], [], [], []
        for _ in range(30):
            st = sim.step()
            steps.append(st.world.pedestrians.current_step)
            p, _ = _ped_arrays(st.world.pedestrians)
            ps.append(p)
            m, v = _det_arrays(st.sensors["pedestrian_detector"].reading, n)
            masks.append(m); vals.append(v)
            colls.append(_mask(st.world.robot_to_pedestrians_collisions, n))
        tag = "loop" if loop else "once"
        res.update({f"{tag}_step": np.array(steps), f"{tag}_poses": np.array(ps), f"{tag}_det_mask": np.array(masks),
                    f"{tag}_det_vals": np.array(vals), f"{tag}_coll_mask": np.array(colls)})
    np.savez(os.path.join(HERE, "replay.npz"), **res)


def gen_eth():
    """ETHPedestriansRecord (pedestrians/_eth.py) on an excerpt of each recording -- the first 14 recorded frames of "eth"
    and frames 40-53 of "hotel" -- at dt = 0.1 (4 simulation frames per recorded gap), 0.05 (8) and 0.03 (13), long
    enough to wrap to frame 0; the policy alone (ids, poses, velocities, frame counters) and under a Simulation with a
    unicycle robot and the detector (readings, collision list).  The reference reads its assets through
    pkg_resources.resource_filename: the excerpt is served by patching that one lookup, the reference's code is unmodified.
    The excerpt's lines are stored in the fixture (they are the input the accelerated class is tested on)."""
    import tempfile
    import pkg_resources
    from pyminisim.pedestrians._eth import ETHPedestriansRecord
    from pyminisim.util import angle_linspace
    res = {}
    real_lookup = pkg_resources.resource_filename
    tmp = tempfile.mkdtemp()
    for seq, first, count in (("eth", 0, 14), ("hotel", 40, 14)):
        lines, frames_seen, last = [], -1, None
        for line in open(real_lookup("pyminisim.pedestrians", f"assets/obsmat_{seq}.txt")):
            fr = int(float(line.split()[0]))
            if fr != last:
                frames_seen += 1
                last = fr
            if first <= frames_seen < first + count:
                lines.append(line)
        with open(os.path.join(tmp, f"obsmat_{seq}.txt"), "w") as f:
            f.writelines(lines)
        res[f"{seq}_lines"] = np.array("".join(lines))
    pkg_resources.resource_filename = lambda pkg, name: os.path.join(tmp, os.path.basename(name))
    try:
        for seq, dt, start, n_steps, robot in (("eth", 0.1, 0, 70, None), ("eth", 0.05, 3, 120, (-6.2, 2.5, 3.0, 0.0, 0.2)),
                                               ("hotel", 0.03, 1, 200, (-1.1, 0.2, 1.5, 0.1, 0.3)), ("hotel", 0.1, 12, 30, None)):
            tag = f"{seq}_{dt}_{start}"
            model = ETHPedestriansRecord(seq, dt, start)
            sim = None
            if robot is not None:
                rm = UnicycleRobotModel(initial_pose=np.array(robot[:3]), initial_control=np.array(robot[3:]))
                sensors = [PedestrianDetector(config=PedestrianDetectorConfig(max_dist=6.0, fov=np.deg2rad(200.0), return_type="absolute"))]
                sim = Simulation(world_map=EmptyWorld(), robot_model=rm, pedestrians_model=model, sensors=sensors, sim_dt=dt, rt_factor=None)
            rows, det, coll = [], [], []        # rows: step, ped id, pose(3), velocity(3); det: step, ped id, value(2); coll: step, ped id
            counters = []

            def record(k, st):
                for ped, pose in st.poses.items():
                    rows.append([k, ped, *pose, *st.velocities[ped]])
                counters.append([st.current_frame, st.current_sub_step])
            record(0, model.state)
            for k in range(1, n_steps + 1):
                if sim is None:
                    model.step(dt, None, None)
                    record(k, model.state)
                else:
                    state = sim.step()
                    record(k, state.world.pedestrians)
                    for ped, v in state.sensors["pedestrian_detector"].reading.pedestrians.items():
                        det.append([k, ped, *v])
                    for ped in state.world.robot_to_pedestrians_collisions:
                        coll.append([k, ped])
            res.update({f"{tag}_rows": np.array(rows), f"{tag}_counters": np.array(counters),
                        f"{tag}_det": np.array(det).reshape(-1, 4), f"{tag}_coll": np.array(coll).reshape(-1, 2)})
    finally:
        pkg_resources.resource_filename = real_lookup
    # util.angle_linspace / wrap_angle known answers (util/_util.py)
    rng = np.random.default_rng(5)
    ang = rng.uniform(-np.pi, np.pi, (40, 2))
    res["angle_in"] = ang
    res["angle_linspace_7"] = np.array([angle_linspace(a, b, 7) for a, b in ang])
    res["wrap_in"] = rng.uniform(-20, 20, 64)
    res["wrap_out"] = wrap_angle(res["wrap_in"])
    np.savez_compressed(os.path.join(HERE, "eth.npz"), **res)


SEMANTIC_OBJECTS = [((3.0, 1.0, 2.0, 1.5), "table", "t1"), ((-2.0, -4.0, 1.0, 3.0), "shelf", None), ((0.5, 3.0, 0.5, 0.5), None, None),
                    ((-4.0, 2.0, 3.0, 0.2), "door", "d7"), ((9.0, 9.0, 1.0, 1.0), "far", "f")]


def gen_semantic():
    """SemanticDetector (sensors/_semantic_object_detector.py) in a Simulation without pedestrians: an AABBWorld with named,
    class-only and anonymous objects, a unicycle robot circling among them; 10 Hz readings without noise and, under
    Simulation.seed(11), with noise (drop_prob 0.4).  Rows: step, object index, nearest point (2)."""
    from pyminisim.sensors import SemanticDetector, SemanticDetectorConfig, SemanticDetectorNoise
    res = {}
    for tag, noise in (("clean", None), ("noisy", SemanticDetectorNoise(drop_prob=0.4))):
        Simulation.seed(11)
        world = AABBWorld([AABBObject(box, class_name=c, object_name=o) for box, c, o in SEMANTIC_OBJECTS])
        rm = UnicycleRobotModel(initial_pose=np.array([0.0, 0.0, 0.5]), initial_control=np.array([1.5, 0.9]))
        sim = Simulation(world_map=world, robot_model=rm, pedestrians_model=None,
                         sensors=[SemanticDetector(SemanticDetectorConfig(max_dist=3.0, frequency=10.), noise=noise)], sim_dt=0.01, rt_factor=None)
        rows, last = [], None
        for k in range(1, 401):
            st = sim.step()
            reading = st.sensors["semantic_detector"].reading
            if reading is not last:                      # a new reading object = the sensor fired on this step
                last = reading
                rows += [[k, idx, *d.position] for idx, d in reading.detections.items()]
                rows.append([k, -1, 0., 0.])             # marks the firing step even when nothing is seen
        res[f"{tag}_rows"] = np.array(rows)
    # AABBWorld.closest_semantic_objects on its own: 60 query points, rows = query index, object index, nearest point
    pts = np.random.default_rng(3).uniform(-6, 6, (60, 2))
    res["query_points"] = pts
    res["query_rows"] = np.array([[q, idx, *v[2]] for q, p in enumerate(pts) for idx, v in world.closest_semantic_objects(p, 2.5).items()])
    np.savez(os.path.join(HERE, "semantic.npz"), **res)


def gen_episode():
    """A 260-step Simulation with every sensor at once (the reference's own scheduler decides who fires when): HSFM crowd of 8,
    unicycle robot with a control sequence driving past obstacles (some moves vetoed by the map), PedestrianDetector
    (period 0), 10 Hz LidarSensor (fires every 11 steps), OmniObstacleDetector (period 0); CirclesWorld and AABBWorld."""
    sc = make_crowd(1, 8, side=5.0, first_world=77)
    circles = np.array([[1.6, 0.9, 0.5], [-2.5, 1.5, 0.7], [0.5, -2.2, 0.6], [3.5, -1.0, 0.4], [-0.5, 3.2, 0.8]])
    boxes = [(2.8, 0.6, 1.2, 1.0), (-1.5, -3.5, 0.8, 1.5), (0.4, 2.6, 0.6, 0.6), (-3.0, 1.0, 1.5, 0.3)]
    K = 260
    ctr = np.stack([np.array([0.9 + 0.3 * np.sin(0.03 * k), 0.8 * np.cos(0.021 * k)]) for k in range(K)])
    res = {"init_poses": sc.poses[0], "table": sc.table[0], "controls": ctr, "circles": circles, "boxes": np.array(boxes),
           "robot_pose0": np.array([-1.0, 0.2, 0.3])}
    for name, world in (("circles", CirclesWorld(circles=circles)), ("aabb", AABBWorld([AABBObject(box=b) for b in boxes]))):
        n = 8
        tracker = FixedWaypointTracker(initial_positions=sc.table[0][:, 0, :].copy(), waypoints=sc.table[0][:, 1:, :].copy(), loop=True)
        model = HeadedSocialForceModelPolicy(waypoint_tracker=tracker, n_pedestrians=n, initial_poses=sc.poses[0].copy(),
                                             pedestrian_linear_velocity_magnitude=1.5)
        rm = UnicycleRobotModel(initial_pose=res["robot_pose0"].copy(), initial_control=np.array([0.0, 0.0]))
        lidar_cfg = LidarSensorConfig(n_beams=37, max_dist=5.0, resolution=0.03)
        sensors = [PedestrianDetector(config=PedestrianDetectorConfig(max_dist=4.0, fov=np.deg2rad(200.0), return_type="relative")),
                   LidarSensor(config=lidar_cfg), OmniObstacleDetector()]
        sim = Simulation(world_map=world, robot_model=rm, pedestrians_model=model, sensors=sensors, sim_dt=0.01, rt_factor=None)
        fired, pts_all, cnt, omni, dmask, dvals, cmask, rpose, rwc, pposes, hits_all = [], [], [], [], [], [], [], [], [], [], []
        prev_lidar = sim.current_state.sensors["lidar"].reading
        for k in range(K):
            st = sim.step(ctr[k])
            rd = st.sensors["lidar"].reading
            if rd is not prev_lidar:            # a new reading object = the scheduler fired (core/_simulation.py:178-180)
                assert st.sensors["lidar"].hold_time == 0.0
                fired.append(k)
                pts = rd.points.reshape(-1, 2)
                pad = np.zeros((lidar_cfg.n_beams, 2)); pad[:len(pts)] = pts
                pts_all.append(pad); cnt.append(len(pts))
                prev_lidar = rd
                # index of the first occupied sample per beam (-1: no hit or the beam starts inside an obstacle): the reference
                # does not return it, so it is re-derived here with the reference's own is_occupied on the very sample
                # points get_reading builds (sensors/_lidar.py:111-119)
                center = st.world.robot.pose[:2]
                angles = np.linspace(0., 2 * np.pi, lidar_cfg.n_beams)
                distances = np.arange(0., lidar_cfg.max_dist, lidar_cfg.resolution)
                cp = np.array([center + np.stack((distances * np.cos(a), distances * np.sin(a)), axis=1) for a in angles])
                occ = world.is_occupied(cp.reshape((-1, 2))).reshape((lidar_cfg.n_beams, -1))
                hit = np.array([int(np.argmax(b)) if (b.any() and not b[0]) else -1 for b in occ], dtype=np.int32)
                assert np.array_equal(cp[hit >= 0, hit[hit >= 0]], pts)
                hits_all.append(hit)
            omni.append(st.sensors["omni_obstacle_detector"].reading.min_obstacle_dist)
            m, v = _det_arrays(st.sensors["pedestrian_detector"].reading, n)
            dmask.append(m); dvals.append(v)
            cmask.append(_mask(st.world.robot_to_pedestrians_collisions, n))
            rpose.append(st.world.robot.pose); rwc.append(int(st.world.robot_to_world_collision))
            pposes.append(_ped_arrays(st.world.pedestrians)[0])
        res.update({f"{name}_lidar_steps": np.array(fired), f"{name}_lidar_points": np.array(pts_all), f"{name}_lidar_count": np.array(cnt), f"{name}_lidar_hit": np.array(hits_all),
                    f"{name}_lidar_hold": np.array(sim.current_state.sensors["lidar"].hold_time),
                    f"{name}_omni": np.array(omni), f"{name}_det_mask": np.array(dmask), f"{name}_det_vals": np.array(dvals),
                    f"{name}_coll_mask": np.array(cmask), f"{name}_robot_pose": np.array(rpose), f"{name}_robot_world_collision": np.array(rwc),
                    f"{name}_ped_poses_last": pposes[-1]})
        print("episode", name, "lidar fires", fired[:4], "...", len(fired), "vetoed robot steps", int(np.sum(rwc)),
              "detections", int(sum(np.count_nonzero(m) for m in dmask)), "lidar points", int(np.sum(cnt)))
    np.savez_compressed(os.path.join(HERE, "episode.npz"), **res)


if __name__ == "__main__":
    which = sys.argv[1:] or ["pair", "kat2", "crowds", "variants", "sensors", "random", "replay", "eth", "semantic", "episode"]
    if "pair" in which:
        gen_pair_forces()
    if "kat2" in which:
        gen_kat2()
    if "crowds" in which:
        gen_crowds()
    if "variants" in which:
        gen_variants()
    if "sensors" in which:
        gen_sensors()
    if "random" in which:
        gen_random_tracker()
    if "replay" in which:
        gen_replay()
    if "eth" in which:
        gen_eth()
    if "semantic" in which:
        gen_semantic()
    if "episode" in which:
        gen_episode()
    print("golden fixtures written to", HERE)
