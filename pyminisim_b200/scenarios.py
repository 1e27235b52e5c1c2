"""Deterministic synthetic inputs for the HSFM / ORCA / sensor hot path.

Pure numpy, no reference and no oracle involved: the same arrays feed the CUDA path, the CPU oracle
and (in the build container) the live reference when golden fixtures are generated.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

# side of the square (metres) at which the reference's explicit-Euler HSFM stays finite for 1000
# steps: denser crowds blow up in the reference itself.
DEFAULT_SIDE = {8: 6.0, 32: 16.0, 64: 40.0}


@dataclass
class CrowdScenario:
    poses: np.ndarray        # (W,N,3) x,y,theta
    vels: np.ndarray         # (W,N,3) vx,vy,omega
    vmag: np.ndarray         # (W,N)
    table: np.ndarray        # (W,N,3,2) FixedWaypointTracker table: row 0 = start, 1 = goal, 2 = start
    robot_pose: np.ndarray   # (W,3)
    robot_control: np.ndarray  # (W,2) unicycle (v, w)
    side: float

    @property
    def n_worlds(self) -> int:
        return self.poses.shape[0]

    @property
    def n_peds(self) -> int:
        return self.poses.shape[1]


def default_side(n_peds: int) -> float:
    if n_peds in DEFAULT_SIDE:
        return DEFAULT_SIDE[n_peds]
    # keep the areal density of the 64-pedestrian case (25 m^2 per pedestrian)
    return float(np.sqrt(25.0 * n_peds))


def make_crowd(n_worlds: int, n_peds: int, side: Optional[float] = None, first_world: int = 0,
               local_goals: Optional[float] = None) -> CrowdScenario:
    """World ``w`` uses ``numpy.random.default_rng(1000 + first_world + w)``.

    Jittered ceil(sqrt(N))^2 grid in a square of side ``side``; heading ~ U(-pi,pi); zero velocity;
    speed 1.5; two looped waypoints per pedestrian: the mirrored start (+ noise) and the start.
    ``local_goals=r`` replaces the mirrored goal with start + U(-r,r)^2 (large single crowds, where
    mirrored goals would funnel everybody through the origin).
    """
    side = default_side(n_peds) if side is None else float(side)
    g = int(np.ceil(np.sqrt(n_peds)))
    cell = side / g
    ix, iy = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    centers = np.stack([(ix.ravel() + 0.5) * cell - side / 2.0, (iy.ravel() + 0.5) * cell - side / 2.0], axis=1)[:n_peds]
    poses = np.zeros((n_worlds, n_peds, 3))
    table = np.zeros((n_worlds, n_peds, 3, 2))
    for w in range(n_worlds):
        rng = np.random.default_rng(1000 + first_world + w)
        start = centers + rng.uniform(-0.2, 0.2, size=(n_peds, 2))
        theta = rng.uniform(-np.pi, np.pi, size=n_peds)
        if local_goals is None:
            goal = -start + rng.uniform(-0.3, 0.3, size=(n_peds, 2))
        else:
            goal = start + rng.uniform(-local_goals, local_goals, size=(n_peds, 2))
        for _ in range(20):  # a goal within 0.35 m of the start gives 0/0 at _hsfm.py:156
            bad = np.linalg.norm(goal - start, axis=1) < 0.35
            if not bad.any():
                break
            if local_goals is None:
                goal[bad] = -start[bad] + rng.uniform(-0.3, 0.3, size=(int(bad.sum()), 2))
            else:
                goal[bad] = start[bad] + rng.uniform(-local_goals, local_goals, size=(int(bad.sum()), 2))
        bad = np.linalg.norm(goal - start, axis=1) < 0.35
        goal[bad] = start[bad] + np.array([1.0, 1.0])
        poses[w, :, :2] = start
        poses[w, :, 2] = theta
        table[w, :, 0] = start
        table[w, :, 1] = goal
        table[w, :, 2] = start
    robot_pose = np.tile(np.array([-side / 2.0 - 1.0, 0.0, 0.0]), (n_worlds, 1))
    robot_control = np.tile(np.array([0.5, 0.0]), (n_worlds, 1))
    return CrowdScenario(poses=poses, vels=np.zeros((n_worlds, n_peds, 3)), vmag=np.full((n_worlds, n_peds), 1.5),
                         table=table, robot_pose=robot_pose, robot_control=robot_control, side=side)


def make_bench_crowd(n_worlds: int, n_peds: int, first_world: int = 0, side: Optional[float] = None) -> CrowdScenario:
    """Benchmark / large-batch variant: goals are start + U(-5,5)^2 and back (looped).

    With thousands of seeds the mirrored-goal crowd drives every world through the origin and the
    reference's explicit-Euler HSFM itself blows up (|v| > 1e5 m/s, then NaN) in 17 % (N = 64) to 39 %
    (N = 32) of the worlds within 1000 steps; with local goals all worlds stay bounded (measured on
    the fp64 CPU restatement of the reference, 256 worlds each; DESIGN.md "workload").
    """
    return make_crowd(n_worlds, n_peds, side=side, first_world=first_world, local_goals=5.0)
