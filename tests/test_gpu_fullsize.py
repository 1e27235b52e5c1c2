"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle needs ~10^4 core-seconds there):
  C4  4096 worlds x 64 pedestrians x 1000 steps: worlds are independent (a world inside the batch == the same world run
      alone, bit for bit), the run is deterministic, sampled worlds agree with the fp64 oracle at a horizon it can afford;
  C3  1 world x 65536 pedestrians: exact pruning == every tile visited, bit for bit; the sorted working copy agrees with the
      caller's-order run to fp32 summation-order noise; forces on a random subsample == direct O(N) sums of the pair force
      (SURVEY.md 8d);
  C5  2048 worlds x 64 ORCA agents: sampled worlds == the same worlds run alone, bit for bit."""
import numpy as np
import pytest

from conftest import rel_err
from helpers import make_cuda, make_oracle
from oracle import oracle as orc
from pyminisim_b200.scenarios import make_bench_crowd, make_crowd

pytestmark = pytest.mark.gpu


def test_c4_full_size_independence_determinism_and_sampled_oracle():
    W, N, K = 4096, 64, 1000
    sc = make_bench_crowd(W, N)
    a = make_cuda(sc, "fp32")
    a.step(0.01, K)
    pa, va = a.get_state()
    out_a = (pa, va) + a.get_detector() + (a.get_collisions(),) + a.get_stats()
    assert np.isfinite(pa).all() and (out_a[-1] == 0).all()
    # deterministic: same bits from a second handle with a different launch configuration
    b = make_cuda(sc, "fp32")
    b.set_tuning(0, 8, 5, 0)
    b.step(0.01, K)
    out_b = b.get_state() + b.get_detector() + (b.get_collisions(),) + b.get_stats()
    for x, y in zip(out_a, out_b):
        assert np.array_equal(x, y)
    # independence: worlds 0, 1234 and 4095 run alone give the bits they have inside the batch -- under the same kernel:
    # a batch this small would be routed to the pair-warp / owner-warp kernel (another summation order, last-bit differences),
    # set_tuning(5) pins the warp-per-world kernel that ran the batch
    assert a.launch_info()["kernel"] == 3
    for w in (0, 1234, 4095):
        one = make_cuda(sc, "fp32", worlds=slice(w, w + 1))
        one.set_tuning(5)
        one.step(0.01, K)
        p1, v1 = one.get_state()
        assert np.array_equal(p1[0], pa[w]) and np.array_equal(v1[0], va[w])
        assert np.array_equal(one.get_detector()[0][0], out_a[2][w])
    # the same property for the pair-warp / owner-warp kernel: 512 worlds (one shard of C4 on 8 GPUs) against worlds run alone
    d = make_cuda(sc, "fp32", worlds=slice(1024, 1536))
    d.step(0.01, K)
    assert d.launch_info()["kernel"] == 4
    pd, vd = d.get_state()
    for w in (0, 300, 511):
        one = make_cuda(sc, "fp32", worlds=slice(1024 + w, 1025 + w))
        one.step(0.01, K)
        assert one.launch_info()["kernel"] == 4
        p1, v1 = one.get_state()
        assert np.array_equal(p1[0], pd[w]) and np.array_equal(v1[0], vd[w])
        assert np.array_equal(one.get_detector()[0][0], d.get_detector()[0][w])
    # and the two kernels agree to the fp32 tolerance of the free-running horizon (test_gpu_trajectory.py states it per world)
    assert rel_err(pd[..., :2], pa[1024:1536, :, :2]) < 1e-4
    # sampled worlds against the fp64 oracle at 300 steps (fp32 tolerance of test_gpu_hsfm.py: 1e-4 @300)
    sel = np.array([7, 511, 2048, 3333])
    sub = make_bench_crowd(W, N)
    for f in ("poses", "vels", "vmag", "table", "robot_pose", "robot_control"):
        setattr(sub, f, getattr(sc, f)[sel])
    o = make_oracle(sub)
    o.rollout(0.01, 300, n_threads=4)
    c = make_cuda(sub, "fp32")
    c.step(0.01, 300)
    assert rel_err(c.get_state()[0], o.poses) < 1e-4


def test_c3_full_size_pruning_exact_and_forces_against_direct_sums():
    n = 65536
    sc = make_crowd(1, n, side=2.0 * np.sqrt(n), local_goals=5.0)
    runs = {}
    for name, tun in (("sorted", (0, 0, 0, 0)), ("pruned", (1, 0, -2, 0)), ("all_tiles", (0, 0, -1, 0))):   # j_split 1: summation order of the cluster kernel
        c = make_cuda(sc, "fp32", robot=0)
        c.set_tuning(*tun)
        c.step(0.01, 3)
        assert c.launch_info()["kernel"] == 1
        runs[name] = c.get_state()
    assert np.array_equal(runs["pruned"][0], runs["all_tiles"][0]) and np.array_equal(runs["pruned"][1], runs["all_tiles"][1])
    assert rel_err(runs["sorted"][0], runs["all_tiles"][0]) < 1e-7       # summation order only
    assert np.abs(runs["sorted"][1] - runs["all_tiles"][1]).max() < 1e-5
    # one-step accelerations of a random subsample against direct sums of the reference's pair force in fp64
    c = make_cuda(sc, "fp32", robot=0)
    c.step(0.01, 1)
    p1, v1 = c.get_state()
    rng = np.random.default_rng(0)
    idx = rng.choice(n, 1024, replace=False)
    pos = sc.poses[0, :, :2]
    m, tau = 70.0, 0.2
    for i in idx[:64]:
        d = pos[i] - pos
        dist = np.linalg.norm(d, axis=1)
        dist[i] = np.inf
        near = dist < 12.0
        f = ((2000.0 * np.exp((0.6 - dist[near]) / 0.08) + 1.2e5 * np.maximum(0.0, 0.6 - dist[near])) / dist[near])[:, None] * d[near]
        fe = f.sum(axis=0)
        goal = sc.table[0, i, 1] - pos[i]
        f0 = m * (goal / np.linalg.norm(goal) * sc.vmag[0, i]) / tau
        th = sc.poses[0, i, 2]
        rf, ro = np.array([np.cos(th), np.sin(th)]), np.array([-np.sin(th), np.cos(th)])
        dvb = np.array([(f0 + fe) @ rf, 1.0 * (fe @ ro)]) / m * 0.01          # v = 0 at t = 0: no k_d term
        th1 = p1[0, i, 2]
        expect = np.array([np.cos(th1) * dvb[0] - np.sin(th1) * dvb[1], np.sin(th1) * dvb[0] + np.cos(th1) * dvb[1]])
        assert np.abs(v1[0, i, :2] - expect).max() < 2e-6 * max(1.0, np.abs(expect).max())


def test_c5_full_size_world_independence():
    from pyminisim_b200.batched import BatchedOrca
    W, N = 2048, 64
    sc = make_crowd(W, N, side=16.0)
    ms = np.full((W, N), 1.7)

    def run(sel):
        c = BatchedOrca(0.01, len(sc.poses[sel]), N)
        c.set_agents(sc.poses[sel][..., :2], max_speeds=ms[sel], pref_speeds=np.full_like(ms[sel], 1.5))
        c.set_waypoints_fixed(sc.table[sel], loop=True)
        c.step(200)
        return c.get_agents()
    pa, va = run(slice(None))
    for w in (0, 1000, 2047):
        p1, v1 = run(slice(w, w + 1))
        assert np.array_equal(p1[0], pa[w]) and np.array_equal(v1[0], va[w])


def test_c4_full_size_sensor_plan_records_are_consistent():
    """C4 with the reference's sensors scheduled inside the ONE launch of 1000 steps (10 Hz lidar, omni, detector masks and
    collision list every step): size-independent properties of the records -- the lidar fires after steps 10, 21, 32, ...
    (core/_simulation.py:177-180), the records' population counts add up to the episode counters the kernel keeps
    independently, the last record equals the readings of the final state, a world inside the batch has the records of the
    same world run alone, and the plan does not change the trajectory."""
    from pyminisim_b200 import _capi as capi
    W, N, K = 4096, 64, 1000
    sc = make_bench_crowd(W, N)
    rng = np.random.default_rng(5)
    circles = np.concatenate([rng.uniform(-25, 25, (W, 8, 2)), rng.uniform(0.3, 1.5, (W, 8, 1))], axis=2)

    def make(worlds=None, plan=True):
        sim = make_cuda(sc, "fp32", worlds=worlds)
        sim.set_map(capi.MAP_CIRCLES, circles if worlds is None else circles[worlds])
        if plan:
            sim.set_lidar(100, 8.0, 0.01)
            sim.schedule_sensor(capi.SENSOR_LIDAR, 0.1)
            sim.schedule_sensor(capi.SENSOR_OMNI, 0.0)
            sim.schedule_sensor(capi.SENSOR_DETECTOR, 0.0)
            sim.schedule_sensor(capi.SENSOR_COLLISIONS, 0.0)
        return sim
    a = make()
    a.step(0.01, K)
    assert a.launch_info()["kernel"] == 3
    steps, hold = a.sensor_fires(capi.SENSOR_LIDAR)
    assert np.array_equal(steps, np.arange(10, K, 11)) and len(steps) == 90
    hit, pts = a.lidar_records()
    omni = a.omni_records()
    dm, cm = a.detector_records(), a.collision_records()
    assert hit.shape == (90, W, 100) and omni.shape == (K, W) and dm.shape == (K, W, 2) and cm.shape == (K, W, 2)
    det, col, nf = a.get_stats()
    pop = lambda m: np.unpackbits(m.view(np.uint8), axis=-1).reshape(m.shape[0], m.shape[1], -1).sum(axis=(0, 2))
    assert np.array_equal(pop(dm), det) and np.array_equal(pop(cm), col) and (nf == 0).all()
    assert np.array_equal(dm[-1], a.get_detector()[0]) and np.array_equal(cm[-1], a.get_collisions())
    assert np.array_equal(omni[-1], a.omni_scan())
    assert (hit >= -1).all() and (hit < 800).all() and (hit >= 0).any()
    sel = hit >= 0                                       # a hit point lies on its beam at hit * resolution from ... some robot position:
    assert np.isfinite(pts[sel]).all() and (pts[~sel] == 0).all()
    # the plan does not touch the dynamics; worlds are independent
    b = make(plan=False)
    b.step(0.01, K)
    assert np.array_equal(a.get_state()[0], b.get_state()[0])
    for w in (0, 2500):
        one = make(worlds=slice(w, w + 1))
        one.set_tuning(5)
        one.step(0.01, K)
        h1, p1 = one.lidar_records()
        assert np.array_equal(h1[:, 0], hit[:, w]) and np.array_equal(p1[:, 0], pts[:, w])
        assert np.array_equal(one.omni_records()[:, 0], omni[:, w]) and np.array_equal(one.detector_records()[:, 0], dm[:, w])
