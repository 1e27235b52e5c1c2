"""ReplayPedestriansPolicy (pedestrians/_replay.py) as a batched gather on the device: one fused launch of recorded frames
under the device robot model and the sensors, against the golden episode of the live reference (tests/golden/replay.npz:
30 Simulation steps, unicycle robot, detector 4 m / 140 deg "relative")."""
import numpy as np
import pytest

from helpers import cuda_available

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not cuda_available(), reason="needs a CUDA device")]


def _sim(g, loop, n_worlds=3):
    from pyminisim_b200 import _capi as capi
    from pyminisim_b200.batched import BatchedSimulation
    sim = BatchedSimulation(n_worlds, 5, precision="fp64")
    sim.set_replay(g["rec_poses"], g["rec_vels"], loop=loop)                      # (T, N, 3): shared by the worlds
    sim.set_robot(capi.ROBOT_UNICYCLE, np.repeat(np.array([[-1.0, 0.2, 0.3]]), n_worlds, 0), control=np.repeat(np.array([[0.8, 0.4]]), n_worlds, 0))
    sim.set_detector(4.0, np.deg2rad(140.0), capi.DET_RELATIVE)
    return sim


@pytest.mark.parametrize("tag,loop", [("once", False), ("loop", True)])
def test_one_launch_of_recorded_frames_matches_reference(golden, tag, loop):
    from pyminisim_b200 import _capi as capi
    g = golden("replay")
    sim = _sim(g, loop)
    sim.schedule_sensor(capi.SENSOR_DETECTOR, 0.0, values=True)
    sim.schedule_sensor(capi.SENSOR_COLLISIONS, 0.0)
    sim.replay_step(0.01, 30)                                                      # ONE launch: 30 frames, robot, sensors
    assert sim.launch_info()["kernel"] == 5
    dm, dv = sim.detector_records(values=True)
    cm = sim.collision_records()
    for w in range(sim.W):
        assert np.array_equal(dm[:, w], g[f"{tag}_det_mask"]) and np.array_equal(cm[:, w], g[f"{tag}_coll_mask"])
        for k in range(30):
            ids = [i for i in range(5) if (int(dm[k, w, 0]) >> i) & 1]
            assert np.allclose(dv[k, w][ids], g[f"{tag}_det_vals"][k][ids], rtol=0, atol=1e-12)
    assert sim.replay_frame() == int(g[f"{tag}_step"][-1])
    poses, _ = sim.get_state()
    assert np.array_equal(poses[0], g[f"{tag}_poses"][-1])                         # the frame shown is the pedestrian state
    mask, vals = sim.get_detector()
    assert np.array_equal(mask, dm[-1]) and np.array_equal(vals, dv[-1])
    # split launches, snapshot / jump
    b = _sim(g, loop, 1)
    b.snapshot_save()
    b.replay_step(0.01, 11)
    b.snapshot_restore()
    assert b.replay_frame() == 0
    for n in (7, 1, 22):
        b.replay_step(0.01, n)
    assert b.replay_frame() == sim.replay_frame()
    assert np.array_equal(b.get_detector()[0][0], mask[0]) and np.array_equal(b.get_collisions()[0], sim.get_collisions()[0])
    assert b.replay_frame(3) == 3 and np.array_equal(b.get_state()[0][0], g["rec_poses"][3])
    sim.close(); b.close()


def test_replay_tick_equals_sense_tick(golden):
    """the drop-in tick (external robot pose in, readings of the next frame out) == pose upload + stand-alone sense pass"""
    g = golden("replay")
    a, b = _sim(g, True, 1), _sim(g, True, 1)
    rng = np.random.default_rng(2)
    T = g["rec_poses"].shape[0]
    for k in range(40):
        rp = np.concatenate([rng.uniform(-2.5, 2.5, 2), rng.uniform(-3, 3, 1)])[None]
        coll, det, vals, frame = a.replay_tick(rp, 0.4)
        assert frame == (k + 1) % T
        c2, d2, v2 = b.sense_tick(g["rec_poses"][frame][None], rp, 0.4)
        assert np.array_equal(coll, c2) and np.array_equal(det, d2) and np.array_equal(vals, v2)
    a.close(); b.close()
