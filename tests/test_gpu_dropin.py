"""Drop-in single-world classes (pyminisim_b200.core / .pedestrians / .sensors) against the golden
fixtures of the live reference: the tests read like the reference's examples."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _ped_arrays(state):
    n = len(state.poses)
    return np.stack([state.poses[i] for i in range(n)]), np.stack([state.velocities[i] for i in range(n)])


def test_kat2_through_simulation(golden):
    from pyminisim_b200.core import Simulation
    from pyminisim_b200.pedestrians import FixedWaypointTracker, HeadedSocialForceModelPolicy
    from pyminisim_b200.robot import UnicycleRobotModel
    from pyminisim_b200.sensors import PedestrianDetector, PedestrianDetectorConfig
    from pyminisim_b200.world_map import EmptyWorld
    g = golden("kat2")
    poses, table = g["init_poses"], g["table"]
    tracker = FixedWaypointTracker(initial_positions=table[:, 0].copy(), waypoints=table[:, 1:].copy(), loop=True)
    model = HeadedSocialForceModelPolicy(waypoint_tracker=tracker, n_pedestrians=3, initial_poses=poses.copy())
    sim = Simulation(world_map=EmptyWorld(), robot_model=UnicycleRobotModel(np.array([0., 0., 0.]), np.array([0.5, 0.2])),
                     pedestrians_model=model, sensors=[PedestrianDetector(PedestrianDetectorConfig(5.0, np.deg2rad(120.), "polar"))],
                     sim_dt=0.01, rt_factor=None)
    k = 0
    for step in range(1, 301):
        st = sim.step()
        assert st.world.robot_to_pedestrians_collisions == []
        if step == g["horizons"][k]:
            p, v = _ped_arrays(st.world.pedestrians)
            assert rel_err(p, g["poses"][k]) < 1e-10
            assert np.allclose(st.world.robot.pose, g["robot_pose"][k], atol=1e-13)
            det = st.sensors["pedestrian_detector"].reading.pedestrians
            ref_ids = [i for i in range(3) if (int(g["det_mask"][k][0]) >> i) & 1]
            assert list(det.keys()) == ref_ids
            for i in ref_ids:
                assert np.allclose(det[i], g["det_vals"][k][i], atol=1e-9)
            k += 1
    # SURVEY.md Appendix C literal
    assert st.sensors["pedestrian_detector"].reading.pedestrians == {}
    # state value semantics: getters return copies
    s = sim.current_state.world.pedestrians
    s.poses[0][0] = 99.0
    assert s.poses[0][0] != 99.0
    # reset_to_state round trip
    snap = sim.current_state.world
    a = sim.step().world.pedestrians.poses[0].copy()
    sim.reset_to_state(snap)
    b = sim.step().world.pedestrians.poses[0]
    assert np.allclose(a, b, atol=1e-12)


def test_kat3_random_tracker_seed_exact(golden):
    """examples/basic.py set-up with Simulation.seed(42): the device pops the waypoint queue, the host replays
    the event log to draw the new tails from the global RNG in the reference's order."""
    from pyminisim_b200.core import Simulation
    from pyminisim_b200.pedestrians import HeadedSocialForceModelPolicy, RandomWaypointTracker
    from pyminisim_b200.robot import UnicycleRobotModel
    from pyminisim_b200.sensors import PedestrianDetector
    from pyminisim_b200.world_map import EmptyWorld
    g = golden("random_tracker")
    Simulation.seed(42)
    robot = UnicycleRobotModel(initial_pose=np.array([0., 0., 0.0]), initial_control=np.array([0.0, np.deg2rad(25.0)]))
    tracker = RandomWaypointTracker(world_size=(7.0, 7.0))
    model = HeadedSocialForceModelPolicy(n_pedestrians=2, waypoint_tracker=tracker,
                                         pedestrian_linear_velocity_magnitude=np.array([1.5, 2.5]),
                                         initial_poses=g["init_poses"].copy())
    sim = Simulation(world_map=EmptyWorld(), robot_model=robot, pedestrians_model=model, sensors=[PedestrianDetector()],
                     rt_factor=None)
    k = 0
    for step in range(1, int(g["horizons"][-1]) + 1):
        st = sim.step()
        if step == g["horizons"][k]:
            p, v = _ped_arrays(st.world.pedestrians)
            cur = np.stack([tracker.state.current_waypoints[i] for i in range(2)])
            assert np.array_equal(cur, g["cur"][k]), step          # waypoints are bit-exact (host RNG replay)
            assert rel_err(p, g["poses"][k]) < 1e-7, (step, rel_err(p, g["poses"][k]))
            k += 1
    assert k == len(g["horizons"])


@pytest.mark.parametrize("world", ["circles", "aabb", "empty"])
def test_lidar_and_omni_sensors_match_reference(golden, world):
    from pyminisim_b200.core import WorldState
    from pyminisim_b200.robot import UnicycleRobotModel
    from pyminisim_b200.sensors import LidarSensor, LidarSensorConfig, LidarSensorNoise, OmniObstacleDetector
    from pyminisim_b200.world_map import AABBObject, AABBWorld, CirclesWorld, EmptyWorld
    g = golden("sensors")
    wm = {"circles": lambda: CirclesWorld(g["circles"]), "aabb": lambda: AABBWorld([AABBObject(box=tuple(b)) for b in g["boxes"]]),
          "empty": EmptyWorld}[world]()
    for cfgname, cfg in (("default", LidarSensorConfig()), ("coarse", LidarSensorConfig(n_beams=37, max_dist=5.0, resolution=0.03))):
        lidar = LidarSensor(cfg)
        assert np.isclose(lidar.period, 0.1) and lidar.sensor_name == "lidar"
        for w, r in enumerate(g["robots"]):
            ws = WorldState(wm.current_state, UnicycleRobotModel(r.copy()).state, None, None, False)
            pts = lidar.get_reading(ws, wm).points.reshape(-1, 2)
            n = int(g[f"lidar_{world}_{cfgname}_count"][w])
            assert len(pts) == n                                               # same beams hit (integer output)
            assert np.array_equal(pts, g[f"lidar_{world}_{cfgname}_points"][w][:n])   # same sample points, bit for bit
    omni = OmniObstacleDetector()
    for w, r in enumerate(g["robots"]):
        ws = WorldState(wm.current_state, UnicycleRobotModel(r.copy()).state, None, None, False)
        d = omni.get_reading(ws, wm).min_obstacle_dist
        assert np.isclose(d, g[f"omni_{world}"][w], rtol=1e-13) or (np.isinf(d) and np.isinf(g[f"omni_{world}"][w]))
    noisy = LidarSensor(LidarSensorConfig(), LidarSensorNoise(drop_prob=0.5))
    ws = WorldState(wm.current_state, UnicycleRobotModel(g["robots"][0].copy()).state, None, None, False)
    assert noisy.get_reading(ws, wm).points.ndim in (1, 2)


def test_batched_lidar_per_world_maps_vs_oracle():
    from oracle import oracle as orc
    from pyminisim_b200.batched import BatchedSimulation
    rng = np.random.default_rng(11)
    W, C = 64, 6
    circles = np.concatenate([rng.uniform(-6, 6, (W, C, 2)), rng.uniform(0.2, 1.2, (W, C, 1))], axis=2)
    robots = np.concatenate([rng.uniform(-3, 3, (W, 2)), rng.uniform(-3, 3, (W, 1))], axis=1)
    boxes = np.concatenate([rng.uniform(-6, 6, (W, C, 2)), rng.uniform(0.3, 2.0, (W, C, 2))], axis=2)
    for kind, data in ((1, circles), (2, boxes)):
        sim = BatchedSimulation(W, 1, precision="fp64")
        sim.set_state(np.zeros((W, 1, 3)))
        sim.set_robot(1, robots, velocity=np.zeros((W, 3)))
        sim.set_map(kind, data)
        o = orc.OracleBatch(np.zeros((W, 1, 3)))
        o.set_robot(orc.ROBOT_EXTERNAL, robots)
        o.set_map(kind, data)
        for cfg in ((100, 8.0, 0.01), (64, 6.0, 0.05)):
            hit, pts = sim.lidar_scan(*cfg)
            ohit, opts = o.lidar(*cfg)
            assert (ohit >= 0).any() and np.array_equal(hit, ohit)             # hit indices: bit-exact
            assert np.array_equal(pts, opts)
        assert np.allclose(sim.omni_scan(), o.omni(), rtol=1e-14)


def test_detector_noise_and_hsfm_noise_paths_run():
    from pyminisim_b200.core import Simulation
    from pyminisim_b200.pedestrians import FixedWaypointTracker, HeadedSocialForceModelPolicy
    from pyminisim_b200.robot import BicycleRobotModel
    from pyminisim_b200.sensors import PedestrianDetector, PedestrianDetectorConfig, PedestrianDetectorNoise
    from pyminisim_b200.world_map import CirclesWorld
    init = np.array([[1.0, 0.2, 0.0], [1.5, -0.4, 1.0], [-2.0, 1.0, 2.0], [0.5, 1.5, -1.0]])
    wps = init[:, None, :2] + np.array([[[3.0, 0.0], [0.0, 3.0]]])
    tracker = FixedWaypointTracker(init[:, :2].copy(), wps, loop=True)
    model = HeadedSocialForceModelPolicy(tracker, 4, initial_poses=init.copy(), noise_std=0.05, precision="fp32")
    det = PedestrianDetector(PedestrianDetectorConfig(6.0, np.deg2rad(300.), "absolute"),
                             PedestrianDetectorNoise(0.0, 0.05, 0.0, 0.01, 0.2))
    Simulation.seed(1)
    sim = Simulation(CirclesWorld(np.array([[8.0, 8.0, 1.0]])), BicycleRobotModel(0.324, np.array([0., 0., 0.]), np.array([0.5, 0.1])),
                     model, [det])
    seen = 0
    for _ in range(50):
        st = sim.step()
        seen += len(st.sensors["pedestrian_detector"].reading.pedestrians)
    assert seen > 0
    p, _ = _ped_arrays(st.world.pedestrians)
    assert np.isfinite(p).all()


def test_orca_dropin_model():
    from oracle import oracle as orc
    from pyminisim_b200.pedestrians import FixedWaypointTracker, ORCAParams, ORCAPedestriansModel
    init = np.array([(4, 0, 0), (0, -4, 0), (0, 4, 0), (2.82, 2.82, 0), (-2.82, -2.82, 0), (2.82, -2.82, 0), (-2.82, 2.82, 0)], dtype=float)
    goals = np.array([[[-4, 0]], [[0, 4]], [[0, -4]], [[-2.82, -2.82]], [[2.82, 2.82]], [[-2.82, 2.82]], [[2.82, -2.82]]], dtype=float)
    ms = np.linspace(1.0, 1.8, 7)
    tracker = FixedWaypointTracker(initial_positions=init[:, :2].copy(), waypoints=goals, loop=True)
    model = ORCAPedestriansModel(0.01, tracker, 7, initial_poses=init.copy(), params=ORCAParams(default_max_speed=2.), max_speeds=ms)
    table = np.concatenate([init[:, None, :2], goals], axis=1)[None]
    o = orc.OracleOrca(0.01, init[None], table, loop=True, max_speed=ms[None], default_max_speed=2.0)
    for _ in range(300):
        model.step()
    o.rollout(300)
    p, v = _ped_arrays(model.state)
    assert np.array_equal(p[:, :2].astype(np.float32), o.pos[0])
    assert np.allclose(p[:, 2], o.heading[0], atol=1e-12)
    assert model.state.waypoints.current_indices == {i: int(o.wp_index[0, i]) for i in range(7)}


@pytest.mark.parametrize("loop", [True, False])
def test_orca_dropin_hand_offs_on_the_array_path(loop):
    """Short legs, so every agent hands over several times in 500 steps: the host mirror of the device tracker (array fast path,
    state object built on demand) against the oracle's tracker -- indices and current waypoints -- and against the dict rule."""
    from oracle import oracle as orc
    from pyminisim_b200.pedestrians import FixedWaypointTracker, ORCAParams, ORCAPedestriansModel
    rng = np.random.default_rng(8)
    n = 9
    init = np.concatenate([rng.uniform(-3, 3, (n, 2)), np.zeros((n, 1))], axis=1)
    goals = init[:, None, :2] + rng.uniform(-1.2, 1.2, (n, 3, 2))
    ms = np.full(n, 1.6)
    tracker = FixedWaypointTracker(initial_positions=init[:, :2].copy(), waypoints=goals.copy(), loop=loop)
    shadow = FixedWaypointTracker(initial_positions=init[:, :2].copy(), waypoints=goals.copy(), loop=loop)   # the dict rule
    model = ORCAPedestriansModel(0.01, tracker, n, initial_poses=init.copy(), params=ORCAParams(default_max_speed=2.), max_speeds=ms)
    table = np.concatenate([init[:, None, :2], goals], axis=1)[None]
    o = orc.OracleOrca(0.01, init[None], table, loop=loop, max_speed=ms[None], default_max_speed=2.0)
    early = None
    for k in range(500):
        model.step()
        shadow.update_waypoints(model.state.poses)
        if k == 120:
            early = model.state                      # evaluated only at the end: must still show step 120's waypoints
            o.rollout(121)
            want_early = {i: table[0, i, int(o.wp_index[0, i])] for i in range(n)}
    o.rollout(500 - 121)
    p, _ = _ped_arrays(model.state)
    assert np.array_equal(p[:, :2].astype(np.float32), o.pos[0])
    idx = {i: int(o.wp_index[0, i]) for i in range(n)}
    assert sum(idx.values()) > n                     # hand-offs happened
    st = model.state.waypoints
    assert st.current_indices == idx == shadow.state.current_indices
    assert all(np.array_equal(st.current_waypoints[i], table[0, i, idx[i]]) for i in range(n))
    assert all(np.array_equal(early.waypoints.current_waypoints[i], want_early[i]) for i in range(n))


def _orca_host_loop(o, tracker, steps):
    """The reference's ORCAPedestriansModel.step around the ORACLE's rvo2 restatement: doStep, tracker hand-off on the new
    poses, preferred velocities from the tracker's current waypoints (_orca.py:95-112,118-127).  The oracle's own table tracker
    is disabled (reach < 0); `tracker` is the drop-in's host-side tracker class."""
    from oracle import oracle as orc
    n = o.N
    for _ in range(steps):
        o.rollout(1)
        poses = np.concatenate([o.pos[0].astype(np.float64), o.heading[0][:, None]], axis=1)
        tracker.update_waypoints({i: poses[i] for i in range(n)})
        cur = tracker.state.current_waypoints
        o.cur_wp[0] = np.stack([np.asarray(cur[i], dtype=np.float64)[:2] for i in range(n)])
        orc.lib().orc_orca_update_pref(__import__("ctypes").byref(o._struct()))


def test_orca_dropin_accepts_random_waypoint_tracker():
    """_orca.py:30-39,106-107 take any AbstractWaypointTracker: with RandomWaypointTracker the hand-off runs on the host (global
    numpy RNG, reference call order) and the device follows the tracker's current waypoints -- bit-identical to the same loop
    around the oracle."""
    from oracle import oracle as orc
    from pyminisim_b200.core import Simulation
    from pyminisim_b200.pedestrians import ORCAParams, ORCAPedestriansModel, RandomWaypointTracker
    n = 9
    rng = np.random.default_rng(5)
    init = np.concatenate([rng.uniform(-3.0, 3.0, (n, 2)), np.zeros((n, 1))], axis=1)
    ms = rng.uniform(1.0, 1.8, n)
    runs = []
    for which in ("cuda", "oracle"):
        Simulation.seed(11)
        tracker = RandomWaypointTracker(world_size=(7.0, 7.0))
        if which == "cuda":
            model = ORCAPedestriansModel(0.01, tracker, n, initial_poses=init.copy(), params=ORCAParams(default_max_speed=2.), max_speeds=ms)
            for _ in range(400):
                model.step()
            p, v = _ped_arrays(model.state)
            runs.append((p[:, :2].astype(np.float32), v[:, :2].astype(np.float32),
                         np.stack([tracker.state.current_waypoints[i] for i in range(n)])))
        else:
            tracker.resample_all({i: init[i] for i in range(n)})
            cur = np.stack([np.asarray(tracker.state.current_waypoints[i], dtype=np.float64)[:2] for i in range(n)])
            table = np.stack([init[:, :2], cur], axis=1)[None]
            o = orc.OracleOrca(0.01, init[None], table, loop=True, max_speed=ms[None], default_max_speed=2.0, reach=-1.0)
            _orca_host_loop(o, tracker, 400)
            runs.append((o.pos[0].copy(), o.vel[0].copy(), np.stack([tracker.state.current_waypoints[i] for i in range(n)])))
    assert np.array_equal(runs[0][2], runs[1][2])            # the same waypoints were drawn and handed off
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])
    assert np.abs(runs[0][0] - init[:, :2].astype(np.float32)).max() > 1.0      # and the agents went somewhere


def test_orca_reset_to_state_reuploads_the_tracker():
    """ORCAPedestriansModel.reset_to_state resets the tracker only (_orca.py:114-116): the next doStep still runs on the old
    preferred velocities, the hand-off and the preferred velocities after it use the reset tracker.  The device copy of the
    tracker must follow (it kept steering to the stale index)."""
    from oracle import oracle as orc
    from pyminisim_b200.pedestrians import (FixedWaypointTracker, FixedWaypointTrackerState, ORCAParams, ORCAPedestriansModel, ORCAState)
    n = 6
    rng = np.random.default_rng(2)
    init = np.concatenate([rng.uniform(-2.5, 2.5, (n, 2)), np.zeros((n, 1))], axis=1)
    goals = rng.uniform(-3.0, 3.0, (n, 3, 2))
    ms = np.full(n, 1.6)
    tracker = FixedWaypointTracker(initial_positions=init[:, :2].copy(), waypoints=goals, loop=True)
    model = ORCAPedestriansModel(0.01, tracker, n, initial_poses=init.copy(), params=ORCAParams(default_max_speed=2.), max_speeds=ms)
    table = np.concatenate([init[:, None, :2], goals], axis=1)[None]
    o = orc.OracleOrca(0.01, init[None], table, loop=True, max_speed=ms[None], default_max_speed=2.0)
    for _ in range(120):
        model.step()
    o.rollout(120)
    # reset: every pedestrian is sent to another row of its table
    new_idx = {i: int((o.wp_index[0, i] + 2) % 4) for i in range(n)}
    new_state = FixedWaypointTrackerState({i: table[0, i, new_idx[i]].copy() for i in range(n)}, dict(new_idx))
    model.reset_to_state(ORCAState({i: (model.state.poses[i], model.state.velocities[i]) for i in range(n)}, new_state))
    for i in range(n):                                        # the oracle: tracker reset, preferred velocities untouched
        o.wp_index[0, i] = new_idx[i]
        o.cur_wp[0, i] = table[0, i, new_idx[i]]
    for _ in range(150):
        model.step()
    o.rollout(150)
    p, v = _ped_arrays(model.state)
    assert np.array_equal(p[:, :2].astype(np.float32), o.pos[0])
    assert model.state.waypoints.current_indices == {i: int(o.wp_index[0, i]) for i in range(n)}


def test_replay_policy_through_simulation(golden):
    """Recorded pedestrians (pedestrians/_replay.py) under Simulation: the device collision list and detector see every frame."""
    from pyminisim_b200.core import Simulation
    from pyminisim_b200.pedestrians import ReplayPedestriansPolicy
    from pyminisim_b200.robot import UnicycleRobotModel
    from pyminisim_b200.sensors import PedestrianDetector, PedestrianDetectorConfig
    from pyminisim_b200.world_map import EmptyWorld
    g = golden("replay")
    for tag, loop in (("once", False), ("loop", True)):
        model = ReplayPedestriansPolicy(g["rec_poses"].copy(), g["rec_vels"].copy(), loop=loop)
        sim = Simulation(world_map=EmptyWorld(), robot_model=UnicycleRobotModel(np.array([-1.0, 0.2, 0.3]), np.array([0.8, 0.4])),
                         pedestrians_model=model, sensors=[PedestrianDetector(PedestrianDetectorConfig(4.0, np.deg2rad(140.0), "relative"))],
                         sim_dt=0.01, rt_factor=None)
        seen = 0
        for k in range(30):
            st = sim.step()
            assert st.world.pedestrians.current_step == g[f"{tag}_step"][k]
            det = st.sensors["pedestrian_detector"].reading.pedestrians
            ref_ids = [i for i in range(5) if (int(g[f"{tag}_det_mask"][k][0]) >> i) & 1]
            assert sorted(det.keys()) == ref_ids
            for i in ref_ids:
                assert np.allclose(det[i], g[f"{tag}_det_vals"][k][i], atol=1e-12)
            assert st.world.robot_to_pedestrians_collisions == [i for i in range(5) if (int(g[f"{tag}_coll_mask"][k][0]) >> i) & 1]
            seen += len(ref_ids)
        assert seen > 30


@pytest.mark.parametrize("visible", [True, False])
def test_fused_tick_equals_separate_sense_pass(visible):
    """Simulation.step takes the collision list and the detector reading from the fused epilogue of the HSFM step
    (pms_step_tick: one call, one synchronisation per tick).  They must equal what a stand-alone sense pass over the
    returned state gives -- the path the reference's own step order describes (core/_simulation.py:138-160)."""
    from pyminisim_b200.core import Simulation, _GpuSenseContext
    from pyminisim_b200.pedestrians import FixedWaypointTracker, HeadedSocialForceModelPolicy
    from pyminisim_b200.robot import UnicycleRobotModel
    from pyminisim_b200.sensors import PedestrianDetector, PedestrianDetectorConfig
    from pyminisim_b200.world_map import EmptyWorld
    from pyminisim_b200 import _capi as capi
    n = 12
    ang = np.linspace(0.0, 2.0 * np.pi, n, endpoint=False)
    start = np.stack([1.7 * np.cos(ang), 1.7 * np.sin(ang)], axis=1)
    poses = np.concatenate([start, (ang + np.pi)[:, None]], axis=1)
    goals = np.stack([-start, start], axis=1)                       # through the robot at the origin and back
    tracker = FixedWaypointTracker(initial_positions=start.copy(), waypoints=goals.copy(), loop=True)
    model = HeadedSocialForceModelPolicy(waypoint_tracker=tracker, n_pedestrians=n, initial_poses=poses.copy(),
                                         robot_visible=visible)
    robot = UnicycleRobotModel(np.array([0.9, 0., 0.]), np.array([0.05, 0.3]))   # in the path of pedestrian 0
    cfg = PedestrianDetectorConfig(5.0, np.deg2rad(120.), "relative")
    sim = Simulation(world_map=EmptyWorld(), robot_model=robot, pedestrians_model=model, sensors=[PedestrianDetector(cfg)],
                     sim_dt=0.01, rt_factor=None)
    check = _GpuSenseContext()
    n_det = n_coll = 0
    for _ in range(250):
        st = sim.step()
        assert model._fused is not None and hasattr(st.world.pedestrians, "_fused_sense")
        ids, coll, det, vals = check.run(st.world.pedestrians.poses, st.world.robot.pose, robot.radius,
                                         (5.0, np.deg2rad(120.), capi.DET_RELATIVE))
        assert st.world.robot_to_pedestrians_collisions == [k for k, c in zip(ids, coll) if c]
        reading = st.sensors["pedestrian_detector"].reading.pedestrians
        assert list(reading.keys()) == [k for k, d in zip(ids, det) if d]
        for k in reading:
            assert reading[k] == (float(vals[k][0]), float(vals[k][1]))
        n_det += len(reading); n_coll += len(st.world.robot_to_pedestrians_collisions)
    assert n_det > 0 and (n_coll > 0 or visible)
