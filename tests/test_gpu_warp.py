"""GPU parity tests of the warp-per-world symmetric-pair kernel (csrc/hsfm_warp.cuh): every tile count
T = ceil(N/32) in 1..4, ragged N (padding lanes), contact-heavy scenes, against the fp64 CPU oracle and
against the directed-pair kernel (csrc/hsfm_small.cuh) on the same input.

Tolerances (max-abs error / max-abs value, conftest.rel_err): fp32 1e-6 @1 step, 1e-5 @100; mixed 1e-7 @1,
1e-5 @100 -- the same bars test_gpu_hsfm.py states for the directed-pair kernel."""
import numpy as np
import pytest

from conftest import rel_err
from helpers import make_oracle
import helpers
from oracle import oracle as orc
from pyminisim_b200.scenarios import make_bench_crowd, make_crowd

pytestmark = pytest.mark.gpu

KERNEL_WARP, KERNEL_SMALL, KERNEL_DUO = 3, 0, 4
# Every test of this file runs twice: on the warp-per-world kernel (j_split 5) and on the pair-warp / owner-warp kernel
# (csrc/hsfm_duo.cuh, j_split 4: FP32 modes, N <= 64; other cases fall back to the warp-per-world kernel).  Without a
# forced choice the library picks by batch size (pms_api.cu:duo_preferred).
FORCE = {"j_split": 5}


@pytest.fixture(autouse=True, params=["warp", "duo"])
def kernel_family(request):
    FORCE["j_split"] = 5 if request.param == "warp" else 4
    yield request.param
    FORCE["j_split"] = 5


def expected_kernel(c):
    duo = FORCE["j_split"] == 4 and c.precision in ("fp32", "fp32_plain") and c.N <= 64
    return KERNEL_DUO if duo else KERNEL_WARP


def make_cuda(sc, precision, **kw):
    c = helpers.make_cuda(sc, precision, **kw)
    c.set_tuning(FORCE["j_split"])
    return c



@pytest.mark.parametrize("precision", ["fp32", "mixed"])
@pytest.mark.parametrize("n", [1, 2, 17, 32, 33, 64, 65, 96, 100, 128])
def test_warp_kernel_vs_oracle(precision, n):
    sc = make_bench_crowd(5, n)
    o = make_oracle(sc)
    c = make_cuda(sc, precision)
    tol = {"fp32": {1: 1e-6, 100: 1e-5}, "mixed": {1: 1e-7, 100: 1e-5}}[precision]
    done = 0
    for h, t in tol.items():
        o.rollout(0.01, h - done, n_threads=4)
        c.step(0.01, h - done)
        done = h
        assert c.launch_info()["kernel"] == expected_kernel(c)
        p, v = c.get_state()
        assert rel_err(p, o.poses) < t, (precision, n, h, rel_err(p, o.poses))
        assert rel_err(v, o.vels) < t * 20, (precision, n, h, rel_err(v, o.vels))
    det, col, nf = c.get_stats()
    assert (nf == 0).all()
    assert (c.get_waypoints()[1] == o.wp_index).all()
    # integer outputs on (nearly) identical state: any mismatch must be a pedestrian sitting on a threshold
    mask, _ = c.get_detector()
    assert np.count_nonzero(mask != o.det_mask) <= 1


@pytest.mark.parametrize("n,side", [(32, 8.0), (64, 11.0), (100, 14.0)])
def test_warp_kernel_dense_contact_scenes(n, side):
    """Mirrored goals in a small square: bodies touch (k_1 / k_2 terms, the rare pass of the kernel)."""
    sc = make_crowd(6, n, side=side)
    o = make_oracle(sc)
    o.rollout(0.01, 120, n_threads=4)
    for precision, tol in (("fp32", 3e-5), ("mixed", 1e-5)):
        c = make_cuda(sc, precision)
        c.step(0.01, 120)
        assert c.launch_info()["kernel"] == expected_kernel(c)
        p, v = c.get_state()
        assert rel_err(p, o.poses) < tol, (precision, n, rel_err(p, o.poses))
    # and the directed-pair kernel agrees to the same level on the same input
    c = make_cuda(sc, "fp32")
    c.set_tuning(-1)
    c.step(0.01, 120)
    assert c.launch_info()["kernel"] == KERNEL_SMALL
    assert rel_err(c.get_state()[0], o.poses) < 3e-5


def test_contact_pass_is_exercised_one_step():
    """Two overlapping pairs in one world, one step: the contact force dominates (k_1 g ~ 1.2e4 N)."""
    from pyminisim_b200.batched import BatchedSimulation
    n = 40
    sc = make_bench_crowd(2, n)
    sc.poses[0, 1, :2] = sc.poses[0, 0, :2] + np.array([0.45, 0.1])      # same tile, overlap 0.14 m
    sc.poses[0, 35, :2] = sc.poses[0, 3, :2] + np.array([-0.2, 0.5])     # across tiles
    sc.vels[0, 0, :2] = [0.7, -0.2]
    sc.vels[0, 35, :2] = [-0.3, 0.9]
    o = make_oracle(sc)
    o.rollout(0.01, 1)
    # g = r_ij - d comes from d = d2 * rsqrt.approx(d2) (2^-22.5 relative): ~6e-7 relative on g = 0.14 m
    for precision, tol in (("fp32", 5e-6), ("mixed", 5e-6)):
        c = make_cuda(sc, precision)
        c.step(0.01, 1)
        p, v = c.get_state()
        assert rel_err(v, o.vels) < tol, (precision, rel_err(v, o.vels))
        assert np.abs(o.vels[0, 1, :2]).max() > 0.5     # the contact really accelerated somebody


@pytest.mark.parametrize("precision", ["fp32", "mixed"])
def test_warp_kernel_tuning_is_bit_identical(precision):
    """Worlds per CTA and the dynamic (chunk, group) scheduler only move work around.  (Pair-warp / owner-warp kernel:
    worlds_per_block 1 .. 5 force its variants -- one / two units of the pair scheme per pair warp, two units compiled for 7 / 4 /
    3 worlds per SM -- bit-identical too.)"""
    sc = make_bench_crowd(37, 64)
    outs = []
    for wpc, chunks in ((0, 1), (1, 1), (2, 1), (3, 1), (4, 1), (5, 1), (8, 4), (7, 3)):
        sim = make_cuda(sc, precision)
        sim.set_tuning(FORCE["j_split"], wpc, chunks)
        sim.step(0.01, 130)
        info = sim.launch_info()
        # (the pair-warp / owner-warp kernel runs one world per CTA on a static grid: wpc / chunks do not apply)
        assert info["kernel"] == expected_kernel(sim) and info["n_chunks"] == (chunks if info["kernel"] == KERNEL_WARP else 1)
        if info["kernel"] == KERNEL_DUO:
            assert info["j_split"] == (wpc if 1 <= wpc <= 5 else 5)           # the variant that ran (37 worlds: the model picks 5)
            assert info["threads_per_block"] == (224 if info["j_split"] == 1 else 160)
            assert info["blocks_per_sm"] == {1: 4, 2: 5, 3: 7, 4: 4, 5: 3}[info["j_split"]]
        outs.append(sim.get_state() + sim.get_detector() + (sim.get_collisions(),) + sim.get_stats() + sim.get_robot())
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b)


def test_warp_kernel_newton_third_law_momentum():
    """No robot, goals switched off (speed 0, tau large): total momentum change comes from pair forces only,
    which cancel exactly in pairs -- the summed velocity stays ~0 while individual velocities do not."""
    from pyminisim_b200.batched import BatchedSimulation, HSFMParams
    n = 64
    sc = make_crowd(4, n, side=9.0)
    sim = BatchedSimulation(4, n, precision="fp32", hsfm_params=HSFMParams(tau=1e9, k_d=0.0, k_o=1.0))
    sim.set_tuning(FORCE["j_split"])
    sim.set_state(sc.poses)
    sim.set_speeds(np.zeros((4, n)))
    sim.set_waypoints_fixed(sc.table, loop=True)
    sim.step(0.01, 1)
    _, v = sim.get_state()
    # body-frame projection keeps only f.r_f and k_o f.r_o = the full force when k_o = 1: v = f dt / m
    tot = np.abs(v[..., :2].sum(axis=1)).max()
    assert np.abs(v[..., :2]).max() > 1e-3
    assert tot < 2e-5 * np.abs(v[..., :2]).sum(axis=1).max() + 1e-6


def _pair(sc, precision, robot_model, det=(5.0, 120.0, 0), loop=True, visible=True, wheel_base=1.0, tracker="fixed"):
    """The same scenario into the oracle and into the CUDA path (robot / tracker configured by the caller)."""
    from pyminisim_b200.batched import BatchedSimulation
    W, N = sc.poses.shape[:2]
    o = orc.OracleBatch(sc.poses, vels=sc.vels, vmag=sc.vmag)
    c = BatchedSimulation(W, N, precision=precision)
    c.set_tuning(FORCE["j_split"])
    c.set_state(sc.poses, sc.vels)
    c.set_speeds(sc.vmag)
    if tracker == "fixed":
        o.set_waypoints_fixed(sc.table, loop=loop)
        c.set_waypoints_fixed(sc.table, loop=loop)
    if robot_model:
        o.set_robot(robot_model, sc.robot_pose, control=sc.robot_control, visible=visible, wheel_base=wheel_base)
        c.set_robot(robot_model, sc.robot_pose, control=sc.robot_control, visible=visible, wheel_base=wheel_base)
    o.set_detector(det[0], np.deg2rad(det[1]), det[2])
    c.set_detector(det[0], np.deg2rad(det[1]), det[2])
    return o, c


def _check(o, c, tol, robot=True):
    p, v = c.get_state()
    assert c.launch_info()["kernel"] == expected_kernel(c)
    assert rel_err(p, o.poses) < tol, rel_err(p, o.poses)
    assert rel_err(v, o.vels) < 20 * tol, rel_err(v, o.vels)
    if robot:
        rp, rv, wc = c.get_robot()
        assert np.allclose(rp, o.robot_pose, atol=1e-11) and np.allclose(rv, o.robot_vel, atol=1e-11)
        assert (wc == o.robot_world_collision).all()


@pytest.mark.parametrize("precision", ["fp32", "mixed"])
def test_warp_kernel_robot_variants(precision):
    """Robot paths of the warp kernel: the unicycle block integrator (with per-step control sequences, several blocks
    and a ragged last block), and the step-by-step path (bicycle, map collision veto, invisible / absent robot)."""
    rng = np.random.default_rng(3)
    sc = make_bench_crowd(5, 40)
    sc.robot_pose[:, :2] = [-4.0, 1.0]
    sc.robot_pose[:, 2] = rng.uniform(-3, 3, 5)
    sc.robot_control[:] = np.stack([rng.uniform(0.2, 0.8, 5), rng.uniform(-0.6, 0.6, 5)], axis=1)
    # 1. unicycle, constant control, 3 calls: 70 steps = 2 blocks + ragged, then 1, then 37
    o, c = _pair(sc, precision, orc.ROBOT_UNICYCLE)
    for n in (70, 1, 37):
        o.rollout(0.01, n, n_threads=4); c.step(0.01, n)
        _check(o, c, 1e-5)
    # 2. unicycle with a per-step control sequence crossing block boundaries; heading wraps through +-pi
    o, c = _pair(sc, precision, orc.ROBOT_UNICYCLE)
    ctl = np.stack([rng.uniform(0.0, 1.0, (90, 5)), rng.uniform(-4.0, 4.0, (90, 5))], axis=-1)
    o.set_controls(ctl); c.set_robot_controls(ctl)
    o.rollout(0.01, 100, n_threads=4); c.step(0.01, 100)     # the last row is held for steps 90..99
    _check(o, c, 1e-5)
    # 3. bicycle (step-by-step path)
    o, c = _pair(sc, precision, orc.ROBOT_BICYCLE, wheel_base=0.324)
    ctl = np.stack([rng.uniform(0.0, 1.0, (60, 5)), rng.uniform(-0.5, 0.5, (60, 5))], axis=-1)
    o.set_controls(ctl); c.set_robot_controls(ctl)
    o.rollout(0.01, 60, n_threads=4); c.step(0.01, 60)
    _check(o, c, 1e-5)
    # 4. unicycle against a circles map: the veto rejects the move (pose and control stay)
    o, c = _pair(sc, precision, orc.ROBOT_UNICYCLE)
    circles = np.array([[-3.0, 1.0, 0.4], [-4.0, 2.5, 0.5], [6.0, 6.0, 1.0]])
    o.set_map(orc.MAP_CIRCLES, circles); c.set_map(1, circles)
    o.rollout(0.01, 150, n_threads=4); c.step(0.01, 150)
    _check(o, c, 1e-5)
    assert o.robot_world_collision.any()
    # 5. invisible robot (sensors still see it) and no robot at all
    o, c = _pair(sc, precision, orc.ROBOT_UNICYCLE, visible=False)
    o.rollout(0.01, 80, n_threads=4); c.step(0.01, 80)
    _check(o, c, 1e-5)
    o, c = _pair(sc, precision, orc.ROBOT_NONE)
    o.rollout(0.01, 80, n_threads=4); c.step(0.01, 80)
    _check(o, c, 1e-5, robot=False)


@pytest.mark.parametrize("precision", ["fp32", "mixed"])
def test_warp_kernel_tracker_variants(precision):
    """Non-loop table (steady freeze, _hsfm.py:302-304), the host-refilled waypoint queue with its event log
    (_random_waypoint_tracker.py:66-91), per-pedestrian speeds, acceleration noise."""
    rng = np.random.default_rng(11)
    sc = make_bench_crowd(4, 70)
    sc.vmag[:] = rng.uniform(0.8, 2.0, sc.vmag.shape)
    # short legs so that waypoints are reached within the horizon
    sc.table[:, :, 1] = sc.table[:, :, 0] + rng.uniform(-1.0, 1.0, sc.table[:, :, 1].shape)
    sc.table[:, :, 1] += np.where(np.linalg.norm(sc.table[:, :, 1] - sc.table[:, :, 0], axis=-1, keepdims=True) < 0.4, 0.6, 0.0)
    # 1. non-loop: pedestrians freeze on the last waypoint
    o, c = _pair(sc, precision, orc.ROBOT_UNICYCLE, loop=False)
    o.rollout(0.01, 260, n_threads=4); c.step(0.01, 260)
    _check(o, c, 3e-5)
    assert (c.get_waypoints()[1] == o.wp_index).all()
    frozen = np.all(o.vels == 0.0, axis=-1)
    assert frozen.any()
    _, v = c.get_state()
    assert (np.all(v == 0.0, axis=-1) == frozen).all()
    # 2. waypoint queue + event log (the queue is long enough not to starve: a pedestrian parked on its last waypoint has an
    #    ill-conditioned desired direction, which no fp32 / fp64 comparison survives)
    W, N = 4, 70
    cur = sc.table[:, :, 1]
    queue = sc.table[:, :, 0][:, :, None, :] + rng.uniform(-1.5, 1.5, (W, N, 8, 2))
    o, c = _pair(sc, precision, orc.ROBOT_UNICYCLE, tracker=None)
    o.set_waypoint_queue(cur, queue); c.set_waypoint_queue(cur, queue)
    o.rollout(0.01, 300, n_threads=4); c.step(0.01, 300)
    _check(o, c, 3e-5)
    ev_c = c.pop_waypoint_events()
    ev_o = o.events_list()
    ev_o = ev_o[np.lexsort((ev_o[:, 2], ev_o[:, 1], ev_o[:, 0]))]
    assert len(ev_c) > 50 and np.array_equal(ev_c, ev_o)
    # 3. acceleration noise (_hsfm.py:288-289) -- drawn by the caller, constant over the call
    o, c = _pair(sc, precision, orc.ROBOT_UNICYCLE)
    noise = rng.normal(0.0, 0.3, (W, N, 2))
    c.set_accel_noise(noise)
    c.step(0.01, 1)
    p1, v1 = c.get_state()
    c2 = _pair(sc, precision, orc.ROBOT_UNICYCLE)[1]
    c2.step(0.01, 1)
    p0, v0 = c2.get_state()
    th = p1[..., 2]                       # heading after the step: v' = v + R(theta') (dv_b + noise) dt
    dv = v1[..., :2] - v0[..., :2]
    exp = 0.01 * np.stack([np.cos(th) * noise[..., 0] - np.sin(th) * noise[..., 1],
                           np.sin(th) * noise[..., 0] + np.cos(th) * noise[..., 1]], axis=-1)
    assert np.abs(dv - exp).max() < 2e-6


def test_async_io_pipeline_equals_blocking_calls():
    """Two handles in asynchronous-I/O mode (copies enqueued, results valid after sync) give the bits of the blocking API."""
    from pyminisim_b200 import _capi as capi
    sc = make_bench_crowd(64, 64)
    ref = make_cuda(sc, "fp32")
    ref.step(0.01, 90)
    p_ref, v_ref = ref.get_state()
    m_ref, d_ref = ref.get_detector()
    det_ref, col_ref, nf_ref = ref.get_stats()
    W, N = 64, 64
    h_p = capi.pinned_array((W, N, 3)); h_p[:] = sc.poses
    h_v = capi.pinned_array((W, N, 3)); h_v[:] = sc.vels
    sims, bufs = [], []
    for _ in range(2):
        s = make_cuda(sc, "fp32")
        s.snapshot_save()
        s.set_async_io(True)
        sims.append(s)
        bufs.append(dict(p=capi.pinned_array((W, N, 3)), v=capi.pinned_array((W, N, 3)),
                         m=capi.pinned_array((W, 2), np.uint32), d=capi.pinned_array((W, N, 2)),
                         det=capi.pinned_array((W,), np.int64), col=capi.pinned_array((W,), np.int64),
                         nf=capi.pinned_array((W,), np.int32)))
    for k in range(4):                       # submit k, then collect k - 1
        s, b = sims[k & 1], bufs[k & 1]
        s.snapshot_restore()
        s.set_state(h_p, h_v)
        s.step_async(0.01, 90)
        s.get_state(b["p"], b["v"])
        s.get_detector(b["m"], b["d"])
        s.get_stats((b["det"], b["col"], b["nf"]))
        if k:
            sims[(k - 1) & 1].sync()
            o = bufs[(k - 1) & 1]
            assert np.array_equal(o["p"], p_ref) and np.array_equal(o["v"], v_ref)
            assert np.array_equal(o["m"], m_ref) and np.array_equal(o["d"], d_ref)
            assert np.array_equal(o["det"], det_ref) and np.array_equal(o["col"], col_ref)
            assert (o["nf"] == 0).all() and (nf_ref == 0).all()   # sync() maps the raw "never" sentinel of the async copy to 0
    sims[1].sync()


def test_automatic_kernel_choice_for_two_tile_batches():
    """pms_api.cu:duo_preferred / duo_upw -- the wave-by-wave model of measured times picks, on a 148-SM
    B200, one of the pair-warp / owner-warp kernel's five variants or the warp-per-world kernel by how the batch fills the SMs."""
    import torch
    from pyminisim_b200.batched import BatchedSimulation
    if torch.cuda.get_device_properties(0).multi_processor_count != 148:
        pytest.skip("the expectations are those of a 148-SM device")
    want = {148: (KERNEL_DUO, 5), 296: (KERNEL_DUO, 1), 444: (KERNEL_DUO, 5), 512: (KERNEL_DUO, 4), 740: (KERNEL_DUO, 2),
            1024: (KERNEL_DUO, 3), 1184: (KERNEL_DUO, 4), 1480: (KERNEL_DUO, 2), 2048: (KERNEL_DUO, 3), 2368: (KERNEL_WARP, 0),
            4096: (KERNEL_WARP, 0)}
    for W, (kernel, variant) in want.items():
        sc = make_bench_crowd(W, 64)
        sim = helpers.make_cuda(sc, "fp32")
        sim.step(0.01, 2)
        info = sim.launch_info()
        assert (info["kernel"], info["j_split"]) == (kernel, variant), (W, info)
        sim.close()
    sim = helpers.make_cuda(make_bench_crowd(1024, 32), "fp32")          # one tile: C2 runs the pair-warp / owner-warp kernel
    sim.step(0.01, 2)
    assert sim.launch_info()["kernel"] == KERNEL_DUO
    sim.close()

