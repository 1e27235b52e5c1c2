"""Device-side noise models: statistical -- not bit -- parity with the reference's numpy draws.

  * HSFM acceleration noise, `noise_std` of HeadedSocialForceModelPolicy (_hsfm.py:208,288-289): N(0, std) per pedestrian
    and body-frame component, drawn afresh in every step (also inside a fused multi-step launch).
  * PedestrianDetectorNoise (sensors/_pedestrian_detector.py:60-72,104-134): misdetection drop, Gaussian distance / angle
    noise in the polar frame, then the frame conversion.

The generator is counter-based (Philox4x32-10 keyed by seed / pedestrian / absolute step), so the tests can also demand
exact reproducibility across launch splits and kernels.  Tolerances: 4 standard errors on means / variances / drop rates,
Kolmogorov-Smirnov p > 1e-3 for normality."""
import numpy as np
import pytest
from scipy import stats

from helpers import make_cuda
from pyminisim_b200.batched import BatchedSimulation
from pyminisim_b200.scenarios import make_bench_crowd

pytestmark = pytest.mark.gpu


def _one_step_noise(sc, precision, std, seed, tuning=None, pre_steps=0):
    """Body-frame noise the kernel applied in one step = R(theta')^T (v_noisy - v_plain) / dt."""
    out = []
    for s in (0.0, std):
        c = make_cuda(sc, precision)
        if tuning:
            c.set_tuning(*tuning)
        if pre_steps:                     # advance the absolute step counter, then put the state back
            c.step(0.01, pre_steps)
            c.set_state(sc.poses, sc.vels)
            c.set_waypoints_fixed(sc.table, loop=True)
            c.set_robot(2, sc.robot_pose, control=sc.robot_control)
        c.set_accel_noise_std(s, seed)
        c.step(0.01, 1)
        out.append(c.get_state())
    (p0, v0), (p1, v1) = out
    assert np.array_equal(p0[..., :2], p1[..., :2])          # positions move with the OLD velocity
    th = p0[..., 2]
    dv = (v1 - v0)[..., :2] / 0.01
    return np.stack([np.cos(th) * dv[..., 0] + np.sin(th) * dv[..., 1], -np.sin(th) * dv[..., 0] + np.cos(th) * dv[..., 1]], -1)


@pytest.mark.parametrize("precision", ["fp32", "mixed", "fp64"])
def test_accel_noise_is_standard_normal_scaled_by_noise_std(precision):
    sc = make_bench_crowd(64, 64)
    std = 0.5
    z = _one_step_noise(sc, precision, std, seed=7) / std
    n = z.size
    assert abs(z.mean()) < 4 / np.sqrt(n)
    assert abs(z.var() - 1.0) < 4 * np.sqrt(2.0 / n) + (2e-3 if precision == "fp32" else 0)
    assert stats.kstest(z.ravel(), "norm").pvalue > 1e-3
    assert abs(np.corrcoef(z[..., 0].ravel(), z[..., 1].ravel())[0, 1]) < 4 / np.sqrt(n / 2)      # x and y independent
    assert abs(np.corrcoef(z[:, :-1].ravel(), z[:, 1:].ravel())[0, 1]) < 4 / np.sqrt(n)            # neighbours independent
    assert abs(stats.kurtosis(z.ravel())) < 0.2


def test_accel_noise_stream_is_keyed_by_seed_pedestrian_and_absolute_step():
    sc = make_bench_crowd(16, 64)
    z0 = _one_step_noise(sc, "fp64", 0.5, seed=7)
    assert np.allclose(_one_step_noise(sc, "fp64", 0.5, seed=7), z0, rtol=0, atol=1e-9)               # reproducible
    assert np.allclose(_one_step_noise(sc, "fp64", 0.25, seed=7), 0.5 * z0, rtol=0, atol=1e-9)        # scales with std
    z_seed = _one_step_noise(sc, "fp64", 0.5, seed=8)
    z_step = _one_step_noise(sc, "fp64", 0.5, seed=7, pre_steps=1)                                    # drawn afresh next step
    for other in (z_seed, z_step):
        assert abs(np.corrcoef(other.ravel(), z0.ravel())[0, 1]) < 0.1
    # every kernel and precision mode sees the same stream: warp kernel (fp32, mixed), directed-pair kernel, large-crowd kernel
    for precision, tuning, tol in (("fp32", None, 2e-3), ("mixed", None, 1e-6), ("fp32", (-1, 0, 0, 0), 2e-3), ("fp64", (0, 0, 0, -4), 1e-6)):
        z = _one_step_noise(sc, precision, 0.5, seed=7, tuning=tuning)
        assert np.abs(z - z0).max() < tol, (precision, tuning, np.abs(z - z0).max())


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_fused_launch_equals_single_step_launches_with_noise(precision):
    """The noise of step t does not depend on how the episode is split into calls or chunks."""
    sc = make_bench_crowd(40, 32)
    outs = []
    for split in ((60,), (1,) * 60, (7, 53), (32, 28)):
        c = make_cuda(sc, precision)
        c.set_accel_noise_std(0.3, 11)
        for k in split:
            c.step(0.01, k)
        outs.append(c.get_state())
    for p, v in outs[1:]:
        assert np.array_equal(p, outs[0][0]) and np.array_equal(v, outs[0][1])
    quiet = make_cuda(sc, precision)
    quiet.step(0.01, 60)
    assert np.abs(quiet.get_state()[1] - outs[0][1]).max() > 1e-3          # and the noise is actually there


def _random_scene(seed=5, W=512, N=48):
    rng = np.random.default_rng(seed)
    poses = np.zeros((W, N, 3))
    poses[..., :2] = rng.uniform(-6, 6, (W, N, 2))
    poses[..., 2] = rng.uniform(-np.pi, np.pi, (W, N))
    robot = np.concatenate([rng.uniform(-2, 2, (W, 2)), rng.uniform(-np.pi, np.pi, (W, 1))], axis=1)
    return poses, robot


def _reading(poses, robot, return_type, noise=None):
    W, N = poses.shape[:2]
    sim = BatchedSimulation(W, N, precision="fp64")
    sim.set_state(poses)
    sim.set_robot(1, robot, velocity=np.zeros((W, 3)))
    sim.set_detector(4.0, np.deg2rad(100.0), return_type)
    if noise is not None:
        sim.set_detector_noise(**noise)
    sim.sense()
    mask, vals = sim.get_detector()
    mask2, vals2 = sim.get_detector()
    assert np.array_equal(mask, mask2) and np.array_equal(vals, vals2)        # one step, one reading
    bits = ((mask[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(W, -1)[:, :N].astype(bool)
    return bits, vals


def test_detector_noise_statistics_and_frame_conversion():
    poses, robot = _random_scene()
    noise = dict(misdetection_prob=0.3, distance_mu=0.2, distance_sigma=0.1, angle_mu=-0.05, angle_sigma=0.04, seed=5)
    exact, ev = _reading(poses, robot, 0)
    noisy, nv = _reading(poses, robot, 0, noise)
    n_det = int(exact.sum())
    assert n_det > 2000
    assert not (noisy & ~exact).any()                                          # noise never invents a detection
    drop = 1.0 - noisy.sum() / n_det
    assert abs(drop - 0.3) < 4 * np.sqrt(0.3 * 0.7 / n_det)
    dd = (nv - ev)[noisy]
    k = len(dd)
    assert abs(dd[:, 0].mean() - 0.2) < 4 * 0.1 / np.sqrt(k) and abs(dd[:, 0].std() - 0.1) < 4 * 0.1 / np.sqrt(2 * k)
    assert abs(dd[:, 1].mean() + 0.05) < 4 * 0.04 / np.sqrt(k) and abs(dd[:, 1].std() - 0.04) < 4 * 0.04 / np.sqrt(2 * k)
    assert stats.kstest((dd[:, 0] - 0.2) / 0.1, "norm").pvalue > 1e-3
    assert stats.kstest((dd[:, 1] + 0.05) / 0.04, "norm").pvalue > 1e-3
    assert abs(np.corrcoef(dd[:, 0], dd[:, 1])[0, 1]) < 4 / np.sqrt(k)
    assert (nv[~noisy] == 0).all()
    # the other return types are the reference's conversion (_pedestrian_detector.py:114-134) of the same noisy polar reading
    dist, ang = nv[..., 0], nv[..., 1]
    xr, yr = dist * np.cos(ang), dist * np.sin(ang)
    th = robot[:, None, 2]
    x, y = xr * np.cos(th) - yr * np.sin(th), xr * np.sin(th) + yr * np.cos(th)
    expect = {1: np.stack([xr, yr], -1), 2: np.stack([x, y], -1), 3: np.stack([x + robot[:, None, 0], y + robot[:, None, 1]], -1)}
    for rt, want in expect.items():
        m, v = _reading(poses, robot, rt, noise)
        assert np.array_equal(m, noisy)
        assert np.allclose(v[noisy], want[noisy], rtol=0, atol=1e-12)
    # another seed: another reading; noise off again: the exact reading
    other, _ = _reading(poses, robot, 0, dict(noise, seed=4))
    assert (other != noisy).any()
    off, ov = _reading(poses, robot, 0, dict(noise, enabled=False))
    assert np.array_equal(off, exact) and np.array_equal(ov, ev)


def test_detector_noise_changes_from_step_to_step_and_leaves_statistics_exact():
    sc = make_bench_crowd(64, 64)
    a = make_cuda(sc, "fp32")
    b = make_cuda(sc, "fp32")
    b.set_detector_noise(misdetection_prob=0.5, distance_sigma=0.1, angle_sigma=0.1, seed=1)
    readings = []
    for _ in range(2):
        a.step(0.01, 150); b.step(0.01, 150)
        ma, va = a.get_detector()
        mb, vb = b.get_detector()
        assert ma.any() and not (mb & ~ma).any() and (mb != ma).any()
        readings.append((ma, mb))
    # the dynamics and the exact episode counters do not see the sensor noise
    assert np.array_equal(a.get_state()[0], b.get_state()[0])
    assert np.array_equal(a.get_stats()[0], b.get_stats()[0])


def test_noise_argument_checks():
    from pyminisim_b200._capi import PmsError
    sim = BatchedSimulation(2, 4)
    with pytest.raises(PmsError):
        sim.set_accel_noise_std(-1.0)
    with pytest.raises(PmsError):
        sim.set_detector_noise(misdetection_prob=1.5)            # assert noise.misdetection_prob <= 1. (:69)
    with pytest.raises(PmsError):
        sim.set_detector_noise(distance_sigma=-0.1)


def test_lidar_and_omni_noise_statistics():
    """LidarSensorNoise (sensors/_lidar.py:52-87) and OmniObstacleDetectorNoise (_omni_obstacle_detector.py:64-69)."""
    rng = np.random.default_rng(11)
    W, C = 256, 6
    circles = np.concatenate([rng.uniform(-6, 6, (W, C, 2)), rng.uniform(0.2, 1.2, (W, C, 1))], axis=2)
    robots = np.concatenate([rng.uniform(-3, 3, (W, 2)), rng.uniform(-3, 3, (W, 1))], axis=1)
    sim = BatchedSimulation(W, 1, precision="fp64")
    sim.set_state(np.zeros((W, 1, 3)))
    sim.set_robot(1, robots, velocity=np.zeros((W, 3)))
    sim.set_map(1, circles)
    hit0, pts0 = sim.lidar_scan(100, 8.0, 0.01)
    d0 = sim.omni_scan()
    cov = np.array([[0.004, 0.0015], [0.0015, 0.002]])
    mean = np.array([0.03, -0.02])
    sim.set_lidar_noise(mean, cov, drop_prob=0.2, seed=9)
    hit1, pts1 = sim.lidar_scan(100, 8.0, 0.01)
    hit2, pts2 = sim.lidar_scan(100, 8.0, 0.01)
    assert np.array_equal(hit1, hit2) and np.array_equal(pts1, pts2)
    was, kept = hit0 >= 0, hit1 >= 0
    n = int(was.sum())
    assert n > 5000 and not (kept & ~was).any() and np.array_equal(hit1[kept], hit0[kept])
    assert abs(1.0 - kept.sum() / n - 0.2) < 4 * np.sqrt(0.2 * 0.8 / n)
    e = (pts1 - pts0)[kept]
    k = len(e)
    assert np.abs(e.mean(axis=0) - mean).max() < 4 * np.sqrt(cov.max() / k)
    assert np.abs(np.cov(e.T) - cov).max() < 6 * cov.max() / np.sqrt(k)
    whitened = np.linalg.solve(np.linalg.cholesky(cov), (e - mean).T).T
    assert stats.kstest(whitened[:, 0], "norm").pvalue > 1e-4 and stats.kstest(whitened[:, 1], "norm").pvalue > 1e-4
    sim.set_lidar_noise(enabled=False)
    assert np.array_equal(sim.lidar_scan(100, 8.0, 0.01)[1], pts0)
    # omni: one reading per world -> gather statistics over seeds
    errs, lost = [], 0
    for seed in range(40):
        sim.set_omni_noise(0.25, 0.1, 0.05, seed=seed)
        d = sim.omni_scan()
        lost += int(np.isinf(d).sum())
        errs.append((d - d0)[np.isfinite(d)])
    errs = np.concatenate(errs)
    total = 40 * W
    assert abs(lost / total - 0.25) < 4 * np.sqrt(0.25 * 0.75 / total)
    assert abs(errs.mean() - 0.1) < 4 * 0.05 / np.sqrt(len(errs)) and abs(errs.std() - 0.05) < 4 * 0.05 / np.sqrt(2 * len(errs))
    assert stats.kstest((errs - 0.1) / 0.05, "norm").pvalue > 1e-4
    from pyminisim_b200._capi import PmsError
    with pytest.raises(PmsError):
        sim.set_lidar_noise(drop_prob=1.5)
    with pytest.raises(PmsError):
        sim.set_lidar_noise(point_cov=np.array([[0.001, 0.01], [0.01, 0.001]]))      # not positive semi-definite
    with pytest.raises(PmsError):
        sim.set_omni_noise(misdetection_prob=2.0)
