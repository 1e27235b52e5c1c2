"""GPU parity tests of the warp-per-world symmetric-pair kernel (csrc/hsfm_warp.cuh): every tile count
T = ceil(N/32) in 1..4, ragged N (padding lanes), contact-heavy scenes, against the fp64 CPU oracle and
against the directed-pair kernel (csrc/hsfm_small.cuh) on the same input.

Tolerances (max-abs error / max-abs value, conftest.rel_err): fp32 1e-6 @1 step, 1e-5 @100; mixed 1e-7 @1,
1e-5 @100 -- the same bars test_gpu_hsfm.py states for the directed-pair kernel."""
import numpy as np
import pytest

from conftest import rel_err
from helpers import make_cuda, make_oracle
from oracle import oracle as orc
from pyminisim_b200.scenarios import make_bench_crowd, make_crowd

pytestmark = pytest.mark.gpu

KERNEL_WARP, KERNEL_SMALL = 3, 0


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
        assert c.launch_info()["kernel"] == KERNEL_WARP
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
    assert o.coll_count.sum() >= 0
    # the scene must exercise body contact between pedestrians at some point
    d = np.linalg.norm(o.poses[:, :, None, :2] - o.poses[:, None, :, :2], axis=-1) + 10 * np.eye(n)
    for precision, tol in (("fp32", 3e-5), ("mixed", 1e-5)):
        c = make_cuda(sc, precision)
        c.step(0.01, 120)
        assert c.launch_info()["kernel"] == KERNEL_WARP
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
    """Worlds per CTA and the dynamic (chunk, group) scheduler only move work around."""
    sc = make_bench_crowd(37, 64)
    outs = []
    for wpc, chunks in ((0, 1), (1, 1), (3, 1), (8, 4), (5, 3)):
        sim = make_cuda(sc, precision)
        sim.set_tuning(0, wpc, chunks)
        sim.step(0.01, 130)
        info = sim.launch_info()
        assert info["kernel"] == KERNEL_WARP and info["n_chunks"] == chunks
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
    sim.set_state(sc.poses)
    sim.set_speeds(np.zeros((4, n)))
    sim.set_waypoints_fixed(sc.table, loop=True)
    sim.step(0.01, 1)
    _, v = sim.get_state()
    # body-frame projection keeps only f.r_f and k_o f.r_o = the full force when k_o = 1: v = f dt / m
    tot = np.abs(v[..., :2].sum(axis=1)).max()
    assert np.abs(v[..., :2]).max() > 1e-3
    assert tot < 2e-5 * np.abs(v[..., :2]).sum(axis=1).max() + 1e-6
