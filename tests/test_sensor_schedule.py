"""Sensors inside a fused launch: the reference's hold-time scheduler (core/_simulation.py:162-185), the lidar / omni
readings of the firing steps and the per-step detector / collision records.

Golden fixture ``episode.npz``: 260 steps of the live reference with every sensor at once (tests/golden/make_golden.py
gen_episode).  CPU: the oracle's episode against the fixture.  GPU: ONE fused launch of 260 steps (and the same episode
split into several launches, and every kernel) against the fixture and against the oracle."""
import numpy as np
import pytest

from oracle import oracle as orc
from helpers import cuda_available

LIDAR = (37, 5.0, 0.03)
DET = (4.0, np.deg2rad(200.0), 2)        # max_dist, fov, return type "relative"
K = 260


def _oracle(g, world, n_worlds=1):
    o = orc.OracleBatch(np.repeat(g["init_poses"][None], n_worlds, 0))
    o.set_waypoints_fixed(np.repeat(g["table"][None], n_worlds, 0), loop=True)
    o.set_robot(orc.ROBOT_UNICYCLE, np.repeat(g["robot_pose0"][None], n_worlds, 0), control=np.zeros((n_worlds, 2)))
    o.set_detector(*DET)
    o.set_map(orc.MAP_CIRCLES if world == "circles" else orc.MAP_AABB, g["circles"] if world == "circles" else g["boxes"])
    return o


def _check_against_golden(g, world, lidar_steps, hit, pts, omni, det_mask, det_vals, coll_mask, w=0, vals_tol=1e-9, exact_points=True):
    """records of world w against the reference's episode"""
    assert np.array_equal(lidar_steps, g[f"{world}_lidar_steps"])           # fires at steps 10, 21, 32, ... (the 11-step quirk)
    assert np.array_equal(hit[:, w], g[f"{world}_lidar_hit"])               # same first occupied sample on every beam (integer output)
    for r in range(len(lidar_steps)):
        sel = hit[r, w] >= 0
        n = int(g[f"{world}_lidar_count"][r])
        assert int(sel.sum()) == n
        # the sample points are robot position + k * resolution * direction: bit-identical when the robot position is (the
        # oracle's robot is; the kernels contract a * b + c of the robot's Euler step into an FMA: last-bit differences)
        if exact_points:
            assert np.array_equal(pts[r, w][sel], g[f"{world}_lidar_points"][r][:n])
        else:
            assert np.allclose(pts[r, w][sel], g[f"{world}_lidar_points"][r][:n], rtol=0, atol=1e-12)
    assert np.allclose(omni[:, w], g[f"{world}_omni"], rtol=1e-12, atol=0)
    assert np.array_equal(det_mask[:, w], g[f"{world}_det_mask"])
    assert np.array_equal(coll_mask[:, w], g[f"{world}_coll_mask"])
    assert np.allclose(det_vals[:, w], g[f"{world}_det_vals"], rtol=0, atol=vals_tol)


@pytest.mark.parametrize("world", ["circles", "aabb"])
def test_oracle_episode_matches_reference(golden, world):
    g = golden("episode")
    o = _oracle(g, world)
    sched = {"detector": [0.0, 0.0], "lidar": [0.1, 0.0], "omni": [0.0, 0.0], "collisions": [0.0, 0.0]}
    ep = o.episode(0.01, K, sched, lidar=LIDAR, controls=g["controls"][:, None, :])
    assert np.array_equal(ep["detector"]["steps"], np.arange(K)) and np.array_equal(ep["omni"]["steps"], np.arange(K))
    _check_against_golden(g, world, ep["lidar"]["steps"], ep["lidar"]["hit"], ep["lidar"]["points"], ep["omni"]["dist"],
                          ep["detector"]["mask"], ep["detector"]["vals"], ep["collisions"]["mask"])
    assert sched["lidar"][1] == float(g[f"{world}_lidar_hold"])                 # the scheduler is left with the same hold_time
    assert np.allclose(o.poses[0], g[f"{world}_ped_poses_last"], rtol=0, atol=1e-9)
    assert np.allclose(o.robot_pose[0], g[f"{world}_robot_pose"][-1], rtol=0, atol=1e-12)


def test_scheduler_eleven_step_quirk_and_carry_over():
    """pure scheduler arithmetic of the oracle helper: 10 Hz at dt = 0.01 fires every 11 steps; hold_time carries over"""
    o = orc.OracleBatch(np.zeros((1, 1, 3)))
    o.set_robot(orc.ROBOT_EXTERNAL, np.zeros((1, 3)))
    s1 = {"omni": [0.1, 0.0]}
    a = o.episode(0.01, 40, s1)["omni"]["steps"]
    assert list(a) == [10, 21, 32]
    s2 = {"omni": [0.1, 0.0]}
    b = list(o.episode(0.01, 15, s2)["omni"]["steps"]) + [15 + t for t in o.episode(0.01, 25, s2)["omni"]["steps"]]
    assert b == [10, 21, 32] and s1["omni"][1] == s2["omni"][1]


# ------------------------------------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu


def _cuda(g, world, precision, n_worlds=1, tuning=None):
    from pyminisim_b200 import _capi as capi
    from pyminisim_b200.batched import BatchedSimulation
    sim = BatchedSimulation(n_worlds, 8, precision=precision)
    sim.set_state(np.repeat(g["init_poses"][None], n_worlds, 0))
    sim.set_waypoints_fixed(np.repeat(g["table"][None], n_worlds, 0), loop=True)
    sim.set_robot(capi.ROBOT_UNICYCLE, np.repeat(g["robot_pose0"][None], n_worlds, 0), control=np.zeros((n_worlds, 2)))
    sim.set_detector(*DET)
    sim.set_map(capi.MAP_CIRCLES if world == "circles" else capi.MAP_AABB, g["circles"] if world == "circles" else g["boxes"])
    sim.set_lidar(*LIDAR)
    sim.schedule_sensor(capi.SENSOR_DETECTOR, 0.0, values=True)
    sim.schedule_sensor(capi.SENSOR_LIDAR, 0.1)
    sim.schedule_sensor(capi.SENSOR_OMNI, 0.0)
    sim.schedule_sensor(capi.SENSOR_COLLISIONS, 0.0)
    if tuning:
        sim.set_tuning(*tuning)
    return sim


def _run(sim, g, n_worlds, splits):
    """the episode in launches of the given lengths; records concatenated"""
    from pyminisim_b200 import _capi as capi
    ctr = np.repeat(g["controls"][:, None, :], n_worlds, 1)
    steps, hit, pts, omni, dm, dv, cm = [], [], [], [], [], [], []
    t0 = 0
    for n in splits:
        sim.set_robot_controls(ctr[t0:t0 + n])
        sim.step(0.01, n)
        s, hold = sim.sensor_fires(capi.SENSOR_LIDAR)
        steps += [t0 + int(t) for t in s]
        h, p = sim.lidar_records()
        hit.append(h); pts.append(p); omni.append(sim.omni_records())
        m, v = sim.detector_records(values=True)
        dm.append(m); dv.append(v); cm.append(sim.collision_records())
        assert np.array_equal(sim.sensor_fires(capi.SENSOR_DETECTOR)[0], np.arange(n))
        t0 += n
    cat = np.concatenate
    return np.array(steps), cat(hit), cat(pts), cat(omni), cat(dm), cat(dv), cat(cm), hold


@gpu
@pytest.mark.skipif(not cuda_available(), reason="needs a CUDA device")
@pytest.mark.parametrize("world", ["circles", "aabb"])
@pytest.mark.parametrize("precision,tuning,kernel", [("fp64", None, 0), ("fp32", (5, 0, 0, 0), 3), ("fp32", (4, 0, 0, 0), 4),
                                                     ("mixed", None, 3), ("fp32", (-1, 0, 0, 0), 0), ("fp64", (0, 0, 0, -4), 1),
                                                     ("fp32", (0, 0, 0, -100), 1)])
def test_fused_launch_returns_the_reference_scans(golden, world, precision, tuning, kernel):
    """ONE launch of 260 fused steps returns exactly the scans the reference emits: lidar at steps 10, 21, 32, ... with the
    reference's hit index on every beam (bit-exact) and its sample points (to 1e-12: the points inherit the last bits of the
    robot position), omni / detector / collision list after every step."""
    g = golden("episode")
    W = 3
    sim = _cuda(g, world, precision, W, tuning)
    steps, hit, pts, omni, dm, dv, cm, hold = _run(sim, g, W, [K])
    assert sim.launch_info()["kernel"] == kernel
    # fp64: the state follows the reference to 1e-9; fp32 state: detector values to fp32 accuracy; the membership masks of this
    # episode have no pedestrian within 1e-5 of a threshold, so they are equal in every precision
    tol = 1e-9 if precision == "fp64" else (2e-6 if precision == "mixed" else 2e-4)
    for w in range(W):
        _check_against_golden(g, world, steps, hit, pts, omni, dm, dv, cm, w=w, vals_tol=tol, exact_points=False)
    assert hold == float(g[f"{world}_lidar_hold"])
    # hit indices (the reference does not expose them) against the oracle's brute-force march
    o = _oracle(g, world)
    ep = o.episode(0.01, K, {"lidar": [0.1, 0.0]}, lidar=LIDAR, controls=g["controls"][:, None, :])
    for w in range(W):
        assert np.array_equal(hit[:, w], ep["lidar"]["hit"][:, 0])
    sim.close()


@gpu
@pytest.mark.skipif(not cuda_available(), reason="needs a CUDA device")
def test_split_episode_fires_at_the_same_steps(golden):
    """hold_time persists on the device: 260 steps as 7 + 30 + 1 + 100 + 122 give the records of one launch, bit for bit;
    a snapshot restores the scheduler with the state"""
    g = golden("episode")
    one = _run(_cuda(g, "circles", "fp64"), g, 1, [K])
    sim = _cuda(g, "circles", "fp64")
    sim.snapshot_save()
    _run(sim, g, 1, [50])
    sim.snapshot_restore()
    parts = _run(sim, g, 1, [7, 30, 1, 100, 122])
    for a, b in zip(one, parts):
        assert np.array_equal(a, b)
    sim.close()


@gpu
@pytest.mark.skipif(not cuda_available(), reason="needs a CUDA device")
def test_records_equal_stand_alone_sensor_calls():
    """records of a fused launch == pms_lidar_scan / pms_omni_scan / pms_get_detector / pms_get_collisions called after
    stepping to the firing step, device noise included (Philox keyed by the absolute step), for a batch of benchmark worlds"""
    from pyminisim_b200 import _capi as capi
    from pyminisim_b200.batched import BatchedSimulation
    from pyminisim_b200.scenarios import make_bench_crowd
    W, N, steps = 40, 64, 45
    sc = make_bench_crowd(W, N)
    rng = np.random.default_rng(3)
    circles = np.concatenate([rng.uniform(-20, 20, (W, 6, 2)), rng.uniform(0.3, 1.5, (W, 6, 1))], axis=2)

    def make():
        sim = BatchedSimulation(W, N, precision="fp32")
        sim.set_state(sc.poses, sc.vels); sim.set_speeds(sc.vmag); sim.set_waypoints_fixed(sc.table, loop=True)
        sim.set_robot(capi.ROBOT_UNICYCLE, sc.robot_pose, control=sc.robot_control)
        sim.set_detector(5.0, np.deg2rad(120.0), capi.DET_POLAR)
        sim.set_map(capi.MAP_CIRCLES, circles)
        sim.set_lidar_noise(drop_prob=0.2, seed=5)
        sim.set_omni_noise(misdetection_prob=0.1, distance_sigma=0.05, seed=6)
        return sim
    a = make()
    a.set_lidar(64, 6.0, 0.05)
    a.schedule_sensor(capi.SENSOR_LIDAR, 0.1); a.schedule_sensor(capi.SENSOR_OMNI, 0.03)
    a.schedule_sensor(capi.SENSOR_DETECTOR, 0.05, values=True); a.schedule_sensor(capi.SENSOR_COLLISIONS, 0.0)
    a.step(0.01, steps)
    fires = {k: a.sensor_fires(k)[0] for k in range(4)}
    assert list(fires[capi.SENSOR_LIDAR]) == [10, 21, 32, 43] and len(fires[capi.SENSOR_COLLISIONS]) == steps
    hit, pts = a.lidar_records(); omni = a.omni_records(); dm, dv = a.detector_records(values=True); cm = a.collision_records()
    b = make()
    for t in range(steps):
        b.step(0.01, 1)
        if t in fires[capi.SENSOR_LIDAR]:
            r = list(fires[capi.SENSOR_LIDAR]).index(t)
            h2, p2 = b.lidar_scan(64, 6.0, 0.05)
            assert np.array_equal(h2, hit[r]) and np.array_equal(p2, pts[r])
        if t in fires[capi.SENSOR_OMNI]:
            assert np.array_equal(b.omni_scan(), omni[list(fires[capi.SENSOR_OMNI]).index(t)])
        if t in fires[capi.SENSOR_DETECTOR]:
            r = list(fires[capi.SENSOR_DETECTOR]).index(t)
            m2, v2 = b.get_detector()
            assert np.array_equal(m2, dm[r]) and np.array_equal(v2, dv[r])
        assert np.array_equal(b.get_collisions(), cm[t])
    a.close(); b.close()
