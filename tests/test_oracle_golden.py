"""Pin the CPU oracle against fixtures generated from the live reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import oracle as orc


def test_pair_force_kat1_and_random(golden):
    g = golden("pair_forces")
    # SURVEY.md Appendix C, KAT-1
    assert np.allclose(g["kat1"][0], [-18980.685914923677, 23999.999999999996], rtol=1e-14)
    assert np.allclose(g["kat1"][1], [-8.08553639890256, -10.780715198536747], rtol=1e-14)
    k = orc.pair_force(0.3, 0.3, [0, 0], [0.5, 0], [1, 0], [0, 1])
    assert np.allclose(k, g["kat1"][0], rtol=1e-13)
    for i in range(len(g["f"])):
        f = orc.pair_force(0.3, float(g["rad_j"][i]), g["r_i"][i], g["r_j"][i], g["v_i"][i], g["v_j"][i])
        assert np.allclose(f, g["f"][i], rtol=1e-12, atol=1e-12 * np.abs(g["f"][i]).max())
    for a, wv in zip(g["angles"], g["wrapped"]):
        assert orc.wrap_angle(a) == wv


def _crowd_batch(poses, table, robot_pose, robot_ctrl, det=(5.0, 120.0, orc.DET_POLAR), loop=True, vmag=1.5,
                 robot=orc.ROBOT_UNICYCLE, visible=True):
    b = orc.OracleBatch(poses, vmag=vmag)
    b.set_waypoints_fixed(table, loop=loop)
    if robot != orc.ROBOT_NONE:
        b.set_robot(robot, robot_pose, control=robot_ctrl, visible=visible, wheel_base=0.324)
    b.set_detector(det[0], np.deg2rad(det[1]), det[2])
    return b


def test_kat2(golden):
    g = golden("kat2")
    b = _crowd_batch(g["init_poses"], g["table"], [0, 0, 0], [0.5, 0.2])
    done = 0
    for k, h in enumerate(g["horizons"]):
        b.rollout(0.01, int(h) - done)
        done = int(h)
        assert rel_err(b.poses[0], g["poses"][k]) < 1e-11
        assert rel_err(b.vels[0], g["vels"][k]) < 1e-10
        assert np.allclose(b.robot_pose[0], g["robot_pose"][k], rtol=0, atol=1e-13)
        assert (b.det_mask[0] == g["det_mask"][k]).all()
        assert (b.coll_mask[0] == g["coll_mask"][k]).all()
        assert np.allclose(b.det_vals[0], g["det_vals"][k], atol=1e-10)
        assert b.det_count[0] == g["det_count"][k]
    # SURVEY.md Appendix C, KAT-2 step 1 and step 100 literals
    assert np.allclose(g["vels"][0][0], [-7.4999993600228479e-02, 1.9251293179769655e-03, 4.9480084294039246], rtol=1e-12)
    assert np.allclose(g["poses"][1][0], [1.115695313189361, 0.6357503710283919, 3.134840772593039], rtol=1e-10)


@pytest.mark.parametrize("n", [8, 32, 64])
def test_crowds(golden, n):
    from pyminisim_b200.scenarios import make_crowd
    g = golden(f"crowd_n{n}")
    W = g["poses"].shape[0]
    sc = make_crowd(W, n)
    b = _crowd_batch(sc.poses, sc.table, sc.robot_pose, sc.robot_control)
    done = 0
    # chaotic amplification of last-bit libm/BLAS differences: 1e-15 -> <=1e-9 @1000
    tol = {1: 1e-13, 100: 1e-11, 300: 1e-10, 1000: 2e-8}
    for k, h in enumerate(g["horizons"]):
        b.rollout(0.01, int(h) - done, n_threads=2)
        done = int(h)
        assert rel_err(b.poses, g["poses"][:, k]) < tol[int(h)], (h, rel_err(b.poses, g["poses"][:, k]))
        assert rel_err(b.vels, g["vels"][:, k]) < tol[int(h)] * 10
        assert (b.det_mask == g["det_mask"][:, k]).all()
        assert (b.coll_mask == g["coll_mask"][:, k]).all()
        assert (b.wp_index == g["wp_index"][:, k]).all()
        assert (b.det_count == g["det_count"][:, k]).all()
        assert (b.coll_count == g["coll_count"][:, k]).all()
    assert (b.nonfinite == 0).all()


def test_variants(golden):
    g = golden("variants")
    poses, table = g["init_poses"], g["table"]
    rp, rc = [-1.0, 0.2, 0.3], [0.4, 0.3]
    hz = [1, 50, 150]

    def run(b, horizons, check, controls=None):
        done = 0
        for k, h in enumerate(horizons):
            if controls is not None:  # controls are per call: one row per step of this call
                b.set_controls(controls[done:h, None, :])
            b.rollout(0.01, h - done)
            done = h
            check(k, b)

    for t, name in enumerate(["polar", "relative_rotated", "relative", "absolute"]):
        b = _crowd_batch(poses, table, rp, rc, det=(4.0, 200.0, t))

        def chk(k, b, name=name):
            assert rel_err(b.poses[0], g[f"rt_{name}_poses"][k]) < 1e-11
            assert (b.det_mask[0] == g[f"rt_{name}_det_mask"][k]).all()
            assert np.allclose(b.det_vals[0], g[f"rt_{name}_det_vals"][k], atol=1e-10)
        run(b, hz, chk)

    b = _crowd_batch(poses, table, rp, rc, visible=False)
    run(b, hz, lambda k, b: (_ for _ in ()).throw(AssertionError) if rel_err(b.poses[0], g["invisible_poses"][k]) > 1e-11 else None)
    b = _crowd_batch(poses, table, None, None, robot=orc.ROBOT_NONE)
    run(b, hz, lambda k, b: (_ for _ in ()).throw(AssertionError) if rel_err(b.vels[0], g["norobot_vels"][k]) > 1e-10 else None)

    # non-loop: freeze at the last waypoint
    b = _crowd_batch(poses, table[:, :2], [-4.0, 0.0, 0.0], [0.3, 0.0], loop=False, vmag=g["freeze_vmag"])

    def chk_freeze(k, b):
        assert rel_err(b.poses[0], g["freeze_poses"][k]) < 1e-10
        assert np.abs(b.vels[0] - g["freeze_vels"][k]).max() < 1e-9
        assert (b.wp_index[0] == g["freeze_wp_index"][k]).all()
        assert (b.det_mask[0] == g["freeze_det_mask"][k]).all()
        assert (b.coll_mask[0] == g["freeze_coll_mask"][k]).all()
    run(b, [1, 200, 400, 700], chk_freeze)
    assert (g["freeze_vels"][-1] == 0).all(axis=1).sum() >= 4  # frozen pedestrians exist in the reference run

    # bicycle + per-step controls
    b = _crowd_batch(poses, table, rp, [0.0, 0.0], robot=orc.ROBOT_BICYCLE)

    def chk_bi(k, b):
        assert np.allclose(b.robot_pose[0], g["bicycle_robot_pose"][k], atol=1e-12)
        assert np.allclose(b.robot_vel[0], g["bicycle_robot_vel"][k], atol=0)
        assert rel_err(b.poses[0], g["bicycle_poses"][k]) < 1e-11
        assert (b.det_mask[0] == g["bicycle_det_mask"][k]).all()
    run(b, hz, chk_bi, g["bicycle_controls"])

    b = _crowd_batch(poses, table, rp, [0.0, 0.0], robot=orc.ROBOT_UNICYCLE)

    def chk_uni(k, b):
        assert np.allclose(b.robot_pose[0], g["unictrl_robot_pose"][k], atol=1e-12)
        assert np.allclose(b.robot_vel[0], g["unictrl_robot_vel"][k], atol=1e-12)
        assert rel_err(b.poses[0], g["unictrl_poses"][k]) < 1e-11
        assert (b.coll_mask[0] == g["unictrl_coll_mask"][k]).all()
    run(b, hz, chk_uni, g["bicycle_controls"])

    # map veto
    b = _crowd_batch(poses, table, [-1.0, 0.2, 0.0], [0.0, 0.0])
    b.set_map(orc.MAP_CIRCLES, g["blocked_circles"])

    def chk_blk(k, b):
        assert np.allclose(b.robot_pose[0], g["blocked_robot_pose"][k], atol=1e-12)
        assert np.allclose(b.robot_vel[0], g["blocked_robot_vel"][k], atol=1e-12)
        assert b.robot_world_collision[0] == g["blocked_robot_world_collision"][k]
        assert rel_err(b.poses[0], g["blocked_poses"][k]) < 1e-11
    run(b, [1, 60, 120, 200], chk_blk, g["blocked_controls"])
    assert g["blocked_robot_world_collision"].any()


def _compact(hit, pts):
    return pts[hit >= 0]


@pytest.mark.parametrize("world", ["circles", "aabb", "empty"])
def test_sensors(golden, world):
    g = golden("sensors")
    robots = g["robots"]
    b = orc.OracleBatch(np.zeros((len(robots), 1, 3)))
    b.set_robot(orc.ROBOT_EXTERNAL, robots)
    if world == "circles":
        b.set_map(orc.MAP_CIRCLES, g["circles"])
    elif world == "aabb":
        b.set_map(orc.MAP_AABB, g["boxes"])
    for cfgname, cfg in (("default", (100, 8.0, 0.01)), ("coarse", (37, 5.0, 0.03))):
        hit, pts = b.lidar(*cfg)
        for w in range(len(robots)):
            ref_n = int(g[f"lidar_{world}_{cfgname}_count"][w])
            mine = _compact(hit[w], pts[w])
            assert len(mine) == ref_n
            assert np.array_equal(mine, g[f"lidar_{world}_{cfgname}_points"][w][:ref_n])
    om = b.omni()
    assert np.allclose(om, g[f"omni_{world}"], rtol=1e-14, equal_nan=True)
    for w in range(len(robots)):
        assert b.map_is_collision(w, robots[w, 0], robots[w, 1], 0.35) == bool(g[f"collision_{world}"][w])
    assert list(g["lidar_fire_steps"]) == [-1, 10, 21, 32]
