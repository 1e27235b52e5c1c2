"""ETHPedestriansRecord (pedestrians/_eth.py) and the util helpers it uses, against fixtures generated from the live reference
(tests/golden/make_golden.py gen_eth: excerpts of the "eth" / "hotel" recordings, the reference's own class run on them).
CPU: the policy alone -- ids, poses, velocities, frame counters, bit for bit.  GPU: the policy under this package's Simulation,
where the interpolated frames live in HBM and a tick is one pms_replay_tick -- detector readings and the collision list."""
import os

import numpy as np
import pytest

from helpers import cuda_available

CASES = [("eth", 0.1, 0), ("eth", 0.05, 3), ("hotel", 0.03, 1), ("hotel", 0.1, 12)]


@pytest.fixture()
def assets(golden, tmp_path):
    g = golden("eth")
    for seq in ("eth", "hotel"):
        (tmp_path / f"obsmat_{seq}.txt").write_text(str(g[f"{seq}_lines"]))
    return str(tmp_path)


def _rows(k, state):
    poses, vels = state.poses, state.velocities
    return [[k, ped, *poses[ped], *vels[ped]] for ped in poses]


def test_util_known_answers(golden):
    from pyminisim_b200.util import angle_linspace, closest_point_of_segment, point_to_segment_distance, wrap_angle
    g = golden("eth")
    assert np.array_equal(wrap_angle(g["wrap_in"]), g["wrap_out"])
    for (a, b), want in zip(g["angle_in"], g["angle_linspace_7"]):
        assert np.array_equal(angle_linspace(a, b, 7), want)
    # geometry: one / many points against one / many segments (util/_geometry.py)
    seg = np.array([[0., 0., 2., 0.], [1., 1., 1., 3.]])
    pts = np.array([[1., 1.], [-1., -1.], [3., 2.]])
    assert np.array_equal(closest_point_of_segment(pts[0], seg[0]), [1., 0.])
    assert np.array_equal(closest_point_of_segment(pts, seg[0]), [[1., 0.], [0., 0.], [2., 0.]])
    assert closest_point_of_segment(pts, seg).shape == (3, 2, 2)
    assert np.allclose(point_to_segment_distance(pts, seg[1]), [0., np.sqrt(8.), 2.])
    with pytest.raises(AssertionError):
        closest_point_of_segment(np.zeros(3), seg[0])


@pytest.mark.parametrize("seq,dt,start", CASES)
def test_policy_alone_matches_the_reference(golden, assets, seq, dt, start):
    from pyminisim_b200.pedestrians import ETHPedestriansRecord
    g = golden("eth")
    tag = f"{seq}_{dt}_{start}"
    want, counters = g[f"{tag}_rows"], g[f"{tag}_counters"]
    model = ETHPedestriansRecord(seq, dt, start, obsmat_path=assets)
    rows = _rows(0, model.state)
    got_counters = [[model.state.current_frame, model.state.current_sub_step]]
    for k in range(1, len(counters)):
        model.step(dt, None, None)
        rows += _rows(k, model.state)
        got_counters.append([model.state.current_frame, model.state.current_sub_step])
    assert np.array_equal(np.array(got_counters), counters)
    assert np.array_equal(np.array(rows), want)               # ids in the recording's order, poses and velocities bit for bit
    assert (np.diff(counters[:, 0]) < 0).any()                # every case runs past the last frame and wraps to frame 0
    # reset_to_state adopts the state object (the counters travel with it)
    snap = model.state
    model.step(dt, None, None)
    after = _rows(0, model.state)
    model.step(dt, None, None)
    model.reset_to_state(snap)
    model.step(dt, None, None)
    assert np.array_equal(np.array(_rows(0, model.state)), np.array(after))


def test_constructor_errors(assets):
    from pyminisim_b200.pedestrians import ETHPedestriansRecord
    with pytest.raises(ValueError, match="Unknown sequence name zara"):
        ETHPedestriansRecord("zara", 0.1, obsmat_path=assets)
    with pytest.raises(AssertionError, match="start_frame must be less than number of frames in sequence which is 14"):
        ETHPedestriansRecord("eth", 0.1, start_frame=14, obsmat_path=assets)
    with pytest.raises(FileNotFoundError):
        ETHPedestriansRecord("eth", 0.1, obsmat_path=os.path.join(assets, "nowhere"))


@pytest.mark.gpu
@pytest.mark.skipif(not cuda_available(), reason="needs a CUDA device")
@pytest.mark.parametrize("seq,dt,start,robot", [("eth", 0.05, 3, (-6.2, 2.5, 3.0, 0.0, 0.2)), ("hotel", 0.03, 1, (-1.1, 0.2, 1.5, 0.1, 0.3))])
def test_under_simulation_frames_on_the_device(golden, assets, seq, dt, start, robot):
    from pyminisim_b200.core import Simulation
    from pyminisim_b200.pedestrians import ETHPedestriansRecord
    from pyminisim_b200.robot import UnicycleRobotModel
    from pyminisim_b200.sensors import PedestrianDetector, PedestrianDetectorConfig
    from pyminisim_b200.world_map import EmptyWorld
    g = golden("eth")
    tag = f"{seq}_{dt}_{start}"
    want, counters, det, coll = g[f"{tag}_rows"], g[f"{tag}_counters"], g[f"{tag}_det"], g[f"{tag}_coll"]
    assert len(det) > 20 and len(coll) > 5
    model = ETHPedestriansRecord(seq, dt, start, obsmat_path=assets)
    rm = UnicycleRobotModel(initial_pose=np.array(robot[:3]), initial_control=np.array(robot[3:]))
    sensors = [PedestrianDetector(config=PedestrianDetectorConfig(max_dist=6.0, fov=np.deg2rad(200.0), return_type="absolute"))]
    sim = Simulation(world_map=EmptyWorld(), robot_model=rm, pedestrians_model=model, sensors=sensors, sim_dt=dt, rt_factor=None)
    rows, got_det, got_coll = _rows(0, model.state), [], []
    kept = None
    for k in range(1, len(counters)):
        if k == 31:                     # jump back to the world of step 20: the device frame counter follows (reset_to_state)
            sim.reset_to_state(kept)
            for j in range(21, 31):
                again = sim.step()
                assert np.array_equal(np.array(_rows(j, again.world.pedestrians)), want[want[:, 0] == j])
        state = sim.step()
        if k == 20:
            kept = state.world
        rows += _rows(k, state.world.pedestrians)
        got_det += [[k, ped, *v] for ped, v in state.sensors["pedestrian_detector"].reading.pedestrians.items()]
        got_coll += [[k, ped] for ped in state.world.robot_to_pedestrians_collisions]
    assert model._sim is not None and model._sim.launch_info()["kernel"] == 5          # the ticks ran as device playback
    assert np.array_equal(np.array(rows), want)
    got_det, got_coll = np.array(got_det).reshape(-1, 4), np.array(got_coll).reshape(-1, 2)
    assert np.array_equal(got_coll, coll)
    assert np.array_equal(got_det[:, :2], det[:, :2])                                   # who is seen, step by step
    # robot pose: the device unicycle and the reference's agree to rounding (FMA contraction), so do absolute readings
    assert np.allclose(got_det[:, 2:], det[:, 2:], rtol=0, atol=1e-12)
