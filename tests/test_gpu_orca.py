"""ORCA kernel against the C++ restatement of the RVO2 agent step (oracle/orca_oracle.cpp).

PARITY UNPINNED against the real rvo2 library (absent from the reference tree and this image; see
oracle/orca_oracle.cpp and DESIGN.md): these tests prove CUDA == restatement, bit for bit in fp32."""
import numpy as np
import pytest

from oracle import oracle as orc
from pyminisim_b200.scenarios import make_crowd

pytestmark = pytest.mark.gpu


def _pair(sc, dt=0.01, loop=True, max_speed=None, pref_speed=None, **kw):
    from pyminisim_b200.batched import BatchedOrca, ORCAParams
    W, N = sc.poses.shape[:2]
    o = orc.OracleOrca(dt, sc.poses, sc.table, loop=loop, max_speed=max_speed, pref_speed=pref_speed, **kw)
    prm = ORCAParams(**{k: v for k, v in kw.items() if k in ("neighbor_dist", "max_neighbors", "time_horizon")})
    c = BatchedOrca(dt, W, N, prm)
    c.set_agents(sc.poses, None, max_speeds=max_speed, pref_speeds=pref_speed)
    c.set_waypoints_fixed(sc.table, loop=loop)
    return o, c


@pytest.mark.parametrize("n,side", [(64, 16.0), (32, 8.0), (7, 3.0)])
def test_orca_matches_restatement_bit_for_bit(n, side):
    sc = make_crowd(12, n, side=side)
    rng = np.random.default_rng(3)
    ms = rng.uniform(1.0, 1.8, (12, n))
    o, c = _pair(sc, max_speed=ms, pref_speed=0.9 * ms)
    done = 0
    seen_constraints = 0
    for h in (1, 50, 400):
        o.rollout(h - done, n_threads=4)
        c.step(h - done)
        done = h
        poses, vels = c.get_agents()
        seen_constraints += int(o.n_constraints.sum())
        assert np.array_equal(poses[..., :2].astype(np.float32), o.pos), h
        assert np.array_equal(vels[..., :2].astype(np.float32), o.vel), h
        assert np.allclose(poses[..., 2], o.heading, atol=1e-12)
        assert (vels[..., 2] == 0).all()
    assert seen_constraints > 0          # the crowd is dense enough to generate ORCA half-planes
    speed = np.linalg.norm(vels[..., :2], axis=-1)
    assert (speed <= ms * (1 + 1e-3) + 1e-3).all()    # the speed disc is respected up to fp32 rounding of far half-planes


def test_orca_examples_orca_py_layout():
    """examples/orca.py: 7 agents on a circle swapping sides (symmetric -> exercises the LP fallbacks)."""
    init = np.array([(4, 0, 0), (0, -4, 0), (0, 4, 0), (2.82, 2.82, 0), (-2.82, -2.82, 0), (2.82, -2.82, 0), (-2.82, 2.82, 0)], dtype=float)
    goals = np.array([[-4, 0], [0, 4], [0, -4], [-2.82, -2.82], [2.82, 2.82], [-2.82, 2.82], [2.82, -2.82]], dtype=float)
    table = np.stack([init[:, :2], goals], axis=1)[None]
    from pyminisim_b200.batched import BatchedOrca, ORCAParams
    ms = np.linspace(1.0, 1.8, 7)[None]
    o = orc.OracleOrca(0.01, init[None], table, loop=True, max_speed=ms, default_max_speed=2.0)
    c = BatchedOrca(0.01, 1, 7, ORCAParams(default_max_speed=2.0))
    c.set_agents(init[None], None, max_speeds=ms)
    c.set_waypoints_fixed(table, loop=True)
    o.rollout(1500)
    c.step(1500)
    poses, vels = c.get_agents()
    assert np.array_equal(poses[..., :2].astype(np.float32), o.pos)
    d = np.linalg.norm(o.pos[0][:, None] - o.pos[0][None], axis=-1) + np.eye(7) * 9
    assert d.min() > 0.55                # nobody interpenetrates (radius 0.3 each)


def test_orca_error_paths():
    from pyminisim_b200.batched import BatchedOrca, ORCAParams
    from pyminisim_b200._capi import PmsError
    c = BatchedOrca(0.01, 1, 4)
    with pytest.raises(PmsError):        # tracker / agents not initialised
        c.step(1)
    with pytest.raises(PmsError):        # _orca.py:57: initial speeds must be <= max_speeds
        c.set_agents(np.zeros((1, 4, 2)), np.full((1, 4, 2), 5.0), max_speeds=np.ones((1, 4)))
    with pytest.raises(PmsError):
        BatchedOrca(0.01, 1, 4, ORCAParams(max_neighbors=99))


def test_orca_async_step_and_snapshot():
    """pms_orca_step_async + pms_sync == pms_orca_step; a snapshot restores positions, velocities and tracker position bit for bit"""
    sc = make_crowd(6, 32, side=8.0)
    _, a = _pair(sc)
    _, b = _pair(sc)
    a.step(40); b.step_async(25); b.step_async(15); b.sync()
    pa, va = a.get_agents(); pb, vb = b.get_agents()
    assert np.array_equal(pa, pb) and np.array_equal(va, vb)
    b.snapshot_save()
    b.step(60)
    b.snapshot_restore()
    b.step(10); a.step(10)
    pa, va = a.get_agents(); pb, vb = b.get_agents()
    assert np.array_equal(pa, pb) and np.array_equal(va, vb)
    a.close(); b.close()
