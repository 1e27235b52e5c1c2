"""Host-side logic of the drop-in classes (no GPU): trackers, maps, robot models, sensor scheduler."""
import numpy as np
import pytest

from pyminisim_b200.core import AbstractSensor, AbstractSensorReading, Simulation
from pyminisim_b200.pedestrians import FixedWaypointTracker, RandomWaypointTracker
from pyminisim_b200.robot import BicycleRobotModel, UnicycleRobotModel
from pyminisim_b200.world_map import AABBObject, AABBWorld, CirclesWorld, EmptyWorld


def test_random_tracker_reproduces_reference_draw_order(golden):
    g = golden("random_tracker")
    Simulation.seed(42)
    tr = RandomWaypointTracker(world_size=(7.0, 7.0))
    init = g["init_poses"]
    tr.resample_all({i: init[i] for i in range(2)})
    cur = np.stack([tr.state.current_waypoints[i] for i in range(2)])
    nxt = np.stack([tr.state.next_waypoints[i] for i in range(2)])
    assert np.array_equal(cur, g["cur0"]) and np.array_equal(nxt, g["next0"])
    # reach -> pop + new tail, everybody else untouched, steady flag always False
    poses = {0: np.append(cur[0], 0.0), 1: np.array([9.0, 9.0, 0.0])}
    out = tr.update_waypoints(poses)
    assert np.array_equal(out[0][0], nxt[0][0]) and out[0][1] is False
    assert np.array_equal(tr.state.next_waypoints[0][:-1], nxt[0][1:])
    assert np.array_equal(tr.state.next_waypoints[1], nxt[1])
    assert np.linalg.norm(tr.state.next_waypoints[0][-1] - nxt[0][-1]) >= 3.0
    pts = tr.sample_independent_points(5, 0.5)
    d = np.linalg.norm(pts[:, None] - pts[None], axis=-1) + np.eye(5)
    assert pts.shape == (5, 2) and d.min() >= 0.5


def test_fixed_tracker_semantics():
    init = np.array([[0.0, 0.0], [5.0, 5.0]])
    wps = np.array([[[1.0, 0.0], [2.0, 0.0]], [[6.0, 5.0], [7.0, 5.0]]])
    tr = FixedWaypointTracker(init, wps, loop=False)
    assert tr.state.current_indices == {0: 1, 1: 1}
    out = tr.update_waypoints({0: np.array([0.9, 0.0, 0.0]), 1: np.array([0.0, 0.0, 0.0])})
    assert list(out) == [0] and out[0][1] is False and np.array_equal(out[0][0], [2.0, 0.0])
    out = tr.update_waypoints({0: np.array([2.0, 0.1, 0.0]), 1: np.array([0.0, 0.0, 0.0])})
    assert out[0][1] is True and tr.state.current_indices[0] == 2          # clamped + steady
    loop = FixedWaypointTracker(init, wps, loop=True)
    loop.update_waypoints({0: np.array([1.0, 0.0, 0.0]), 1: np.array([6.0, 5.0, 0.0])})
    out = loop.update_waypoints({0: np.array([2.0, 0.0, 0.0]), 1: np.array([0.0, 0.0, 0.0])})
    assert loop.state.current_indices[0] == 0 and np.array_equal(out[0][0], init[0])
    ids, table = loop.table()
    assert ids == [0, 1] and table.shape == (2, 3, 2) and np.array_equal(table[:, 0], init)


def test_maps_match_reference(golden):
    g = golden("sensors")
    robots = g["robots"]
    worlds = {"circles": CirclesWorld(g["circles"]), "aabb": AABBWorld([AABBObject(box=tuple(b)) for b in g["boxes"]]),
              "empty": EmptyWorld()}
    for name, wm in worlds.items():
        for w, r in enumerate(robots):
            assert bool(wm.is_collision(r[:2].copy(), 0.35)) == bool(g[f"collision_{name}"][w])
            d = wm.closest_distance_to_obstacle(r[:2].copy())
            assert np.isclose(d, g[f"omni_{name}"][w], rtol=1e-13) or (np.isinf(d) and np.isinf(g[f"omni_{name}"][w]))
    pts = np.random.default_rng(0).uniform(-5, 5, (200, 2))
    occ = worlds["circles"].is_occupied(pts)
    ref = (np.linalg.norm(pts[:, None] - g["circles"][None, :, :2], axis=-1) - g["circles"][None, :, 2] <= 0).any(axis=1)
    assert np.array_equal(occ, ref)


def test_robot_models_match_reference(golden):
    g = golden("variants")
    world = CirclesWorld(g["blocked_circles"])
    rob = UnicycleRobotModel(np.array([-1.0, 0.2, 0.0]), np.array([0.0, 0.0]))
    ctr = g["blocked_controls"]
    blocked = 0
    for k, h in enumerate([1, 60, 120, 200]):
        while blocked < h:
            ok = rob.step(0.01, world, ctr[min(blocked, len(ctr) - 1)])
            blocked += 1
        assert np.allclose(rob.state.pose, g["blocked_robot_pose"][k], atol=1e-13)
        assert np.allclose(rob.state.velocity, g["blocked_robot_vel"][k], atol=1e-13)
        assert (not ok) == bool(g["blocked_robot_world_collision"][k])
    bike = BicycleRobotModel(0.324, np.array([-1.0, 0.2, 0.3]), np.array([0.0, 0.0]))
    done = 0
    for k, h in enumerate([1, 50, 150]):
        while done < h:
            bike.step(0.01, EmptyWorld(), g["bicycle_controls"][done])
            done += 1
        assert np.allclose(bike.state.pose, g["bicycle_robot_pose"][k], atol=1e-13)
        assert (bike.state.velocity == 0).all()


def test_sensor_scheduler_fires_every_eleven_steps(golden):
    """hold_time accumulates 0.01 in fp64: a 10 Hz sensor fires at 0-based steps 10, 21, 32."""
    fired = []
    box = [-1]

    class Dummy(AbstractSensor):
        def __init__(self):
            super().__init__("lidar", period=1. / 10.)

        sensor_config = property(lambda self: None)

        def get_reading(self, world_state, world_map):
            fired.append(box[0])
            return AbstractSensorReading.__new__(type("R", (AbstractSensorReading,), {}))

    sim = Simulation(EmptyWorld(), UnicycleRobotModel(np.array([0., 0., 0.])), None, [Dummy()], sim_dt=0.01)
    for k in range(40):
        box[0] = k
        sim.step()
    assert fired == list(golden("sensors")["lidar_fire_steps"]) == [-1, 10, 21, 32]
    # enable / disable semantics (core/_simulation.py:85-117)
    sim.set_robot_enabled(False)
    assert sim.current_state.world.robot is None and sim.current_state.sensors == {}
    sim.set_robot_enabled(True)
    import pytest
    with pytest.raises(RuntimeError):
        sim.set_pedestrians_enabled(False)


def test_replay_policy_frame_sequence_matches_reference(golden):
    """ReplayPedestriansPolicy (pedestrians/_replay.py:46-56): frame index and poses per step, loop and no-loop."""
    from pyminisim_b200.pedestrians import ReplayPedestriansPolicy, ReplayPedestriansState
    g = golden("replay")
    for tag, loop in (("once", False), ("loop", True)):
        model = ReplayPedestriansPolicy(g["rec_poses"], g["rec_vels"], loop=loop)
        assert model.state.current_step == 0
        for k in range(30):
            model.step(0.01, None, None)
            assert model.state.current_step == g[f"{tag}_step"][k]
            assert np.array_equal(np.stack([model.state.poses[i] for i in range(5)]), g[f"{tag}_poses"][k])
        snap = model.state
        model.step()
        model.reset_to_state(snap)
        assert model.state.current_step == snap.current_step
    with pytest.raises(AssertionError):
        ReplayPedestriansPolicy(np.zeros((4, 3, 2)), np.zeros((4, 3, 3)))
    with pytest.raises(AssertionError):
        ReplayPedestriansState(-1, {})


def test_fixed_tracker_array_path_equals_the_dict_rule():
    """FixedWaypointTracker.update_from_array (the accelerated models' host mirror: positions as an (N, 2) array, no per-pedestrian
    objects unless somebody reached a waypoint) against update_waypoints (_fixed_waypoint_tracker.py:68-87) on a random walk that
    crosses many waypoints, loop and clamp; state_later() shows the waypoints of the step it was taken at."""
    from pyminisim_b200.pedestrians import FixedWaypointTracker
    rng = np.random.default_rng(12)
    n = 11
    start = rng.uniform(-2, 2, (n, 2))
    wps = start[:, None, :] + rng.uniform(-1.0, 1.0, (n, 4, 2))
    for loop in (True, False):
        a = FixedWaypointTracker(start.copy(), wps.copy(), reach_distance=0.35, loop=loop)
        b = FixedWaypointTracker(start.copy(), wps.copy(), reach_distance=0.35, loop=loop)
        pos = start.copy()
        snaps = []
        for step in range(400):
            # walk towards the current waypoint of tracker b with some noise
            cur = np.stack([b.state.current_waypoints[i] for i in range(n)])
            pos = pos + 0.08 * (cur - pos) / np.maximum(np.linalg.norm(cur - pos, axis=1, keepdims=True), 1e-9) + rng.normal(0, 0.01, (n, 2))
            idx = a.update_from_array(pos)
            b.update_waypoints({i: np.array([pos[i, 0], pos[i, 1], 0.0]) for i in range(n)})
            assert a.state.current_indices == b.state.current_indices == {i: int(idx[i]) for i in range(n)}
            if step % 97 == 0:
                snaps.append((a.state_later(), {i: b.state.current_waypoints[i].copy() for i in range(n)}))
        assert sum(b.state.current_indices.values()) > n + 5                     # hand-offs happened
        for later, want in snaps:
            got = later().current_waypoints
            assert all(np.array_equal(got[i], want[i]) for i in range(n))
        # the dict path and the array path share one state: mixing them stays coherent
        a.update_waypoints({i: np.array([pos[i, 0], pos[i, 1], 0.0]) for i in range(n)})
        assert np.array_equal(a.update_from_array(pos), np.array([a.state.current_indices[i] for i in range(n)]))

