"""Drop-in mirror of ``pyminisim.pedestrians`` for the accelerated path.

``HeadedSocialForceModelPolicy`` and ``ORCAPedestriansModel`` keep the reference's constructor signature
and ``state`` / ``step`` / ``reset_to_state`` contract (pedestrians/_hsfm.py:196-313, _orca.py:28-127) but
every step runs in libpms_b200 on the GPU (one world = a batch of one).  The waypoint trackers hold the
host-side state the reference exposes; the reach test, index hand-off and queue pop happen on the device.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

import numpy as np

from . import _capi as capi
from .batched import BatchedOrca, BatchedSimulation, HSFMParams, ORCAParams
from .util import angle_linspace, wrap_angle
from .core import (DEFAULT_ROBOT_RADIUS, PEDESTRIAN_RADIUS, AbstractPedestriansModel, AbstractPedestriansModelState,
                   AbstractWaypointTracker, AbstractWaypointTrackerState)

__all__ = ["HSFMParams", "HeadedSocialForceModelPolicy", "RandomWaypointTracker", "FixedWaypointTracker",
           "ORCAParams", "ORCAPedestriansModel", "HSFMState", "ORCAState", "ReplayPedestriansPolicy", "ReplayPedestriansState",
           "ETHPedestriansRecord", "ETHState"]


# ---------------------------------------------------------------------------------------------------------
# FixedWaypointTracker (pedestrians/_fixed_waypoint_tracker.py)
# ---------------------------------------------------------------------------------------------------------
class FixedWaypointTrackerState(AbstractWaypointTrackerState):
    def __init__(self, current_waypoints: Dict[int, np.ndarray], current_indices: Optional[Dict[int, int]] = None):
        if current_indices is None:
            current_indices = {k: 0 for k in current_waypoints.keys()}
        else:
            assert current_waypoints.keys() == current_indices.keys()
        super().__init__(current_waypoints)
        self._current_indices = current_indices

    @property
    def current_indices(self) -> Dict[int, int]:
        return self._current_indices


class FixedWaypointTracker(AbstractWaypointTracker):
    def __init__(self, initial_positions, waypoints, reach_distance=0.3, loop: bool = False):
        self._reach_distance, self._loop = reach_distance, loop
        if isinstance(waypoints, np.ndarray):
            assert isinstance(initial_positions, np.ndarray), "Initial positions must have same type as waypoints"
            assert initial_positions.ndim == 2 and initial_positions.shape[1] == 2, \
                f"Initial positions must have shape of (n_pedestrians, 2) the {initial_positions.shape} array is given"
            assert initial_positions.shape[0] == waypoints.shape[0], "Initial positions and waypoints must have same number of pedestrians"
            assert waypoints.ndim == 3 and waypoints.shape[2] == 2, "Waypoints must have shape of (n_pedestrians, n_points, 2)"
            ids = range(waypoints.shape[0])
        else:
            assert isinstance(initial_positions, dict), "Initial positions must have same type as waypoints"
            for k, v in waypoints.items():
                assert k in initial_positions, f"Pedestrian {k} doesn't have initial position"
                assert initial_positions[k].shape == (2,)
                assert v.ndim == 2 and v.shape[1] == 2
            ids = waypoints.keys()
        # row 0 of every table is the initial position; tracking starts at row 1 (:45,59)
        self._all_waypoints = {k: np.concatenate((np.asarray(initial_positions[k], dtype=float)[None, :],
                                                   np.asarray(waypoints[k], dtype=float)), axis=0) for k in ids}
        self._current_indices = {k: 1 for k in self._all_waypoints}
        rows = {v.shape[0] for v in self._all_waypoints.values()}
        self._stacked = (list(self._all_waypoints.keys()), np.stack(list(self._all_waypoints.values()))) if len(rows) == 1 else None
        self._idx_cache = None

    # -- device hand-off helpers --
    @property
    def reach_distance(self):
        return self._reach_distance

    @property
    def loop(self):
        return self._loop

    def table(self) -> Tuple[List[int], np.ndarray]:
        ids = list(self._all_waypoints.keys())
        rows = {self._all_waypoints[k].shape[0] for k in ids}
        assert len(rows) == 1, "the accelerated path needs the same number of waypoints for every pedestrian"
        return ids, np.stack([self._all_waypoints[k] for k in ids])

    def _set_indices(self, idx):
        self._current_indices = {k: int(i) for k, i in zip(self._all_waypoints.keys(), idx)}
        self._idx_cache = None

    @property
    def state(self) -> FixedWaypointTrackerState:
        return FixedWaypointTrackerState(self._current(), self._current_indices)

    def resample_all(self, agents_poses):
        return self._current()

    def update_waypoints(self, agents_poses):
        assert agents_poses.keys() == self._all_waypoints.keys()
        out = {}
        self._idx_cache = None
        if self._stacked is not None:      # same rule, one distance evaluation for all pedestrians
            keys, tabs = self._stacked
            idx = np.fromiter((self._current_indices[k] for k in keys), dtype=np.int64, count=len(keys))
            pos = np.stack([agents_poses[k][:2] for k in keys])
            reached = np.sqrt(((tabs[np.arange(len(keys)), idx] - pos) ** 2).sum(axis=1)) <= self._reach_distance
            for j in np.nonzero(reached)[0]:
                k, nxt, steady = keys[j], int(idx[j]) + 1, False
                if nxt >= tabs.shape[1]:
                    nxt, steady = (0, False) if self._loop else (tabs.shape[1] - 1, True)
                self._current_indices[k] = nxt
                out[k] = (self._all_waypoints[k][nxt], steady)
            return out
        for k, pose in agents_poses.items():
            tab, idx = self._all_waypoints[k], self._current_indices[k]
            if np.linalg.norm(tab[idx] - pose[:2]) <= self._reach_distance:
                nxt, steady = idx + 1, False
                if nxt >= tab.shape[0]:
                    nxt, steady = (0, False) if self._loop else (tab.shape[0] - 1, True)
                self._current_indices[k] = nxt
                out[k] = (tab[nxt], steady)
        return out

    def update_from_array(self, positions: np.ndarray) -> np.ndarray:
        """update_waypoints for the accelerated models, whose pedestrians are rows 0..N-1 of an array: same rule, no per-pedestrian
        Python objects unless somebody reached a waypoint.  Returns the (N,) table indices after the hand-off."""
        keys, tabs = self._stacked
        idx = self._index_array()
        reached = np.sqrt(((tabs[self._rows, idx] - positions) ** 2).sum(axis=1)) <= self._reach_distance
        if reached.any():
            for j in np.nonzero(reached)[0]:
                nxt = int(idx[j]) + 1
                if nxt >= tabs.shape[1]:
                    nxt = 0 if self._loop else tabs.shape[1] - 1
                idx[j] = nxt
                self._current_indices[keys[j]] = nxt
        return idx

    def _index_array(self) -> np.ndarray:
        # cache of the indices dict as an array (the dict stays the tracker's state: states alias it like in the reference)
        cache = self._idx_cache
        if cache is None or cache[0] is not self._current_indices:
            keys = self._stacked[0]
            self._rows = np.arange(len(keys))
            cache = self._idx_cache = (self._current_indices, np.fromiter((self._current_indices[k] for k in keys), dtype=np.int64, count=len(keys)))
        return cache[1]

    def state_later(self):
        """The state as of now, built when somebody asks for it: current waypoints from a copy of the indices taken now, the
        indices dict the tracker's own (the reference's state holds that very dict, _fixed_waypoint_tracker.py:62-66)."""
        keys, tabs = self._stacked
        idx, live = self._index_array().copy(), self._current_indices
        return lambda: FixedWaypointTrackerState({k: tabs[j, idx[j], :] for j, k in enumerate(keys)}, live)

    def reset_to_state(self, state: FixedWaypointTrackerState):
        self._current_indices = state.current_indices

    def sample_independent_points(self, n_points: int, min_cross_distance: Optional[float] = None):
        pass

    def _current(self):
        return {k: self._all_waypoints[k][i, :] for k, i in self._current_indices.items()}


# ---------------------------------------------------------------------------------------------------------
# RandomWaypointTracker (pedestrians/_random_waypoint_tracker.py): global numpy RNG, same draw order
# ---------------------------------------------------------------------------------------------------------
class RandomWaypointTrackerState(AbstractWaypointTrackerState):
    def __init__(self, current_waypoints: Dict[int, np.ndarray], next_waypoints: Dict[int, np.ndarray]):
        super().__init__(current_waypoints)
        self._next_waypoints = next_waypoints

    @property
    def next_waypoints(self) -> Dict[int, np.ndarray]:
        return self._next_waypoints


class RandomWaypointTracker(AbstractWaypointTracker):
    def __init__(self, world_size: Tuple[float, float], min_sample_distance=3.0, max_sample_trials=100,
                 reach_distance=0.3, n_next_waypoints: int = 10):
        assert len(world_size) == 2
        assert n_next_waypoints >= 1
        self._world_size = world_size
        self._min_sample_distance, self._max_sample_trials = min_sample_distance, max_sample_trials
        self._reach_distance, self._n_next_waypoints = reach_distance, n_next_waypoints
        self._state: Optional[RandomWaypointTrackerState] = None

    @property
    def reach_distance(self):
        return self._reach_distance

    @property
    def min_sample_distance(self):
        return self._min_sample_distance

    @property
    def state(self) -> RandomWaypointTrackerState:
        return self._state

    def _draw(self) -> np.ndarray:
        half = self._world_size[0] / 2.       # both axes use world_size[0], like the reference (:101-102,112-113)
        return np.random.uniform(low=np.array([-half, -half]), high=np.array([half, half]))

    def sample_waypoint(self, others: np.ndarray, min_distance: float) -> np.ndarray:
        """Rejection sampling against one point or a set of points (_random_waypoint_tracker.py:110-123)."""
        for _ in range(self._max_sample_trials):
            cand = self._draw()
            if others.ndim == 1:
                dist = np.linalg.norm(others - cand)
            elif others.ndim == 2:
                dist = np.min(np.linalg.norm(others - cand, axis=1))
            else:
                raise RuntimeError("Unsupported agents positions shape")
            if dist < min_distance:
                continue
            return cand
        raise RuntimeError("Failed to sample waypoint")

    def resample_all(self, agents_poses):
        keys = sorted(agents_poses.keys())
        start = np.stack([agents_poses[k] for k in keys], axis=0)[:, :2]
        chain = np.zeros((len(keys), self._n_next_waypoints + 1, 2))
        chain[:, 0] = np.stack([self.sample_waypoint(p, self._min_sample_distance) for p in start])
        for j in range(1, self._n_next_waypoints + 1):     # column by column: every pedestrian, then the next column
            chain[:, j] = np.stack([self.sample_waypoint(p, self._min_sample_distance) for p in chain[:, j - 1]])
        cur = {k: chain[n, 0] for n, k in enumerate(keys)}
        nxt = {k: chain[n, 1:] for n, k in enumerate(keys)}
        self._state = RandomWaypointTrackerState(cur, nxt)
        return cur

    def update_waypoints(self, agents_poses):
        if self._state is None:
            raise RuntimeError("Waypoint tracker is not initialized")
        assert agents_poses.keys() == self._state.current_waypoints.keys()
        cur, nxt = self._state.current_waypoints, self._state.next_waypoints
        new_cur, new_nxt = {}, {}
        for k, pose in agents_poses.items():
            if np.linalg.norm(cur[k] - pose[:2]) <= self._reach_distance:
                new_cur[k], new_nxt[k] = self.advance(nxt[k])
            else:
                new_cur[k], new_nxt[k] = cur[k].copy(), nxt[k].copy()
        self._state = RandomWaypointTrackerState(new_cur, new_nxt)
        return {k: (v.copy(), False) for k, v in new_cur.items()}

    def advance(self, queue: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """Pop the queue front and draw a new tail >= min_sample_distance from the old tail (:83-87)."""
        shifted = np.zeros_like(queue)
        shifted[:-1] = queue[1:]
        shifted[-1] = self.sample_waypoint(queue[-1], self._min_sample_distance)
        return queue[0], shifted

    def reset_to_state(self, state: RandomWaypointTrackerState):
        self._state = state

    def sample_independent_points(self, n_points: int, min_cross_distance: Optional[float] = None):
        assert n_points > 0
        if min_cross_distance is None:
            min_cross_distance = -np.inf
        pts = self._draw().reshape(1, -1)
        for _ in range(1, n_points):
            pts = np.vstack([pts, self.sample_waypoint(pts, min_cross_distance).reshape(1, -1)])
        return pts


# ---------------------------------------------------------------------------------------------------------
# HSFM
# ---------------------------------------------------------------------------------------------------------
class _ArrayBackedState(AbstractPedestriansModelState):
    """State object of the accelerated models.  Takes the reference's signature -- a dict ``{id: (pose, velocity)}``
    (core/_pedestrians_model.py:12-19) -- or, from the models themselves, a pair of ``(N, 3)`` arrays that the state owns.
    In the second case nothing per-pedestrian is built until somebody asks: ``poses`` / ``velocities`` return a fresh dict of
    copies on every access like the reference (:21-27), ``pose_array`` / ``velocity_array`` hand out the arrays without the dict."""

    def __init__(self, pedestrians, waypoints_state):
        if isinstance(pedestrians, dict):
            self._arrays = None
            super().__init__(pedestrians, waypoints_state)
        else:
            poses, velocities = pedestrians
            assert poses.ndim == 2 and poses.shape[1] == 3 and velocities.shape == poses.shape
            self._arrays = (poses, velocities)
            self._dict = None
            self._waypoints = waypoints_state        # a tracker state, or a zero-argument callable that builds it on first use

    @property
    def waypoints(self):
        if callable(self._waypoints):
            self._waypoints = self._waypoints()
        return self._waypoints

    @property
    def _pedestrians(self):                       # what the base class (and code written against it) reads
        if self._dict is None:
            p, v = self._arrays
            self._dict = {i: (p[i], v[i]) for i in range(p.shape[0])}
        return self._dict

    @_pedestrians.setter
    def _pedestrians(self, value):
        self._dict = value

    @property
    def poses(self):
        if self._arrays is None:
            return {k: v[0].copy() for k, v in self._dict.items()}
        p = self._arrays[0].copy()
        return {i: p[i] for i in range(p.shape[0])}

    @property
    def velocities(self):
        if self._arrays is None:
            return {k: v[1].copy() for k, v in self._dict.items()}
        v = self._arrays[1].copy()
        return {i: v[i] for i in range(v.shape[0])}

    @property
    def pose_array(self) -> np.ndarray:
        """(N, 3) copy of the poses in pedestrian-id order (ids of the accelerated models are 0..N-1)."""
        if self._arrays is None:
            return np.stack([self._dict[k][0] for k in sorted(self._dict)])
        return self._arrays[0].copy()

    @property
    def velocity_array(self) -> np.ndarray:
        if self._arrays is None:
            return np.stack([self._dict[k][1] for k in sorted(self._dict)])
        return self._arrays[1].copy()


class HSFMState(_ArrayBackedState):
    pass


def _mask_ids(words: np.ndarray, n: int):
    """Pedestrian ids whose bit is set in a (ceil(n/32),) uint32 mask, ascending (the reference's list order)."""
    return np.flatnonzero(np.unpackbits(np.ascontiguousarray(words).view(np.uint8), bitorder="little")[:n]).tolist()


def _initial_arrays(tracker, n, initial_poses, initial_velocities):
    if initial_poses is None:   # _hsfm.py:210-214 / _orca.py:41-45: positions from the tracker, headings ~ U(-pi, pi)
        positions = tracker.sample_independent_points(n, 0.5)
        headings = np.random.uniform(-np.pi, np.pi, size=n)
        initial_poses = np.hstack([positions, headings.reshape(-1, 1)])
    else:
        assert initial_poses.shape[0] == n
    if initial_velocities is None:
        initial_velocities = np.zeros((n, 3))
    else:
        assert initial_velocities.shape[0] == n
    return np.asarray(initial_poses, dtype=np.float64), np.asarray(initial_velocities, dtype=np.float64)


class _TrackerLink:
    """Keeps a host tracker object and the device-side tracker state of one world in sync."""

    def __init__(self, tracker: AbstractWaypointTracker, n: int):
        self.tracker, self.n = tracker, n
        self.random = isinstance(tracker, RandomWaypointTracker)
        if not self.random and not isinstance(tracker, FixedWaypointTracker):
            raise TypeError("the accelerated path supports FixedWaypointTracker and RandomWaypointTracker")

    def upload(self, sim):
        st = self.tracker.state
        if self.random:
            cur = np.stack([st.current_waypoints[i] for i in range(self.n)])[None]
            queue = np.stack([st.next_waypoints[i] for i in range(self.n)])[None]
            sim.set_waypoint_queue(cur, queue, reach_distance=self.tracker.reach_distance)
        else:
            ids, table = self.tracker.table()
            assert ids == list(range(self.n)), "pedestrian ids must be 0..n-1"
            idx = np.array([st.current_indices[i] for i in range(self.n)], dtype=np.int32)[None]
            sim.set_waypoints_fixed(table[None], start_index=idx, reach_distance=self.tracker.reach_distance, loop=self.tracker.loop)

    def after_step(self, sim, tick=None):
        """``tick`` (BatchedSimulation.step_tick): event count and waypoint indices arrived with the state already."""
        if self.random:
            if tick is not None and int(tick["n_events"][0]) == 0:
                return
            events = sim.pop_waypoint_events()
            if len(events) == 0:
                return
            st = self.tracker.state
            cur, nxt = dict(st.current_waypoints), dict(st.next_waypoints)
            appended_ids, appended_pts = [], []
            for step, _, ped in events:          # (step, world, ped) order == the reference's draw order
                if step < 0:
                    raise RuntimeError("waypoint queue starved on the device (too many steps per call)")
                cur[ped], nxt[ped] = self.tracker.advance(nxt[ped])
                appended_ids.append((0, int(ped)))
                appended_pts.append(nxt[ped][-1])
            self.tracker.reset_to_state(RandomWaypointTrackerState(cur, nxt))
            sim.append_waypoints(np.array(appended_ids, dtype=np.int32), np.array(appended_pts))
        else:
            idx = tick["wp_index"] if tick is not None else sim.get_waypoints()[1]
            self.tracker._set_indices(idx[0])


class HeadedSocialForceModelPolicy(AbstractPedestriansModel):
    """GPU-backed HSFM with the constructor of pedestrians/_hsfm.py:198-208.  ``precision`` selects the
    arithmetic mode of the kernel ("fp64" reproduces the reference to ~1e-15 per step)."""

    def __init__(self, waypoint_tracker: AbstractWaypointTracker, n_pedestrians: int,
                 initial_poses: Optional[np.ndarray] = None, initial_velocities: Optional[np.ndarray] = None,
                 hsfm_params: HSFMParams = HSFMParams.create_default(), pedestrian_mass: float = 70.0,
                 pedestrian_linear_velocity_magnitude: Union[float, np.ndarray] = 1.5, robot_visible: bool = True,
                 robot_radius: float = DEFAULT_ROBOT_RADIUS, noise_std: Optional[float] = None,
                 precision: str = "fp64", device: int = 0):
        poses, vels = _initial_arrays(waypoint_tracker, n_pedestrians, initial_poses, initial_velocities)
        self._n = poses.shape[0]
        if isinstance(pedestrian_linear_velocity_magnitude, np.ndarray):
            assert pedestrian_linear_velocity_magnitude.shape == (n_pedestrians,), \
                "Linear velocity magnitude must be float or (n_pedestrians,) shape ndarray"
            vmag = pedestrian_linear_velocity_magnitude.astype(np.float64)
        else:
            vmag = np.repeat(float(pedestrian_linear_velocity_magnitude), self._n)
        self._tracker = waypoint_tracker
        if self._tracker.state is None:
            self._tracker.resample_all({i: poses[i] for i in range(self._n)})
        self._robot_visible, self._robot_radius, self._noise_std = robot_visible, robot_radius, noise_std
        self._sim = BatchedSimulation(1, self._n, precision=precision, device=device, hsfm_params=hsfm_params,
                                      pedestrian_mass=pedestrian_mass, pedestrian_radius=PEDESTRIAN_RADIUS,
                                      hsfm_robot_radius=robot_radius)
        self._sim.set_speeds(vmag[None])
        self._sim.set_state(poses[None], vels[None])
        self._link = _TrackerLink(self._tracker, self._n)
        self._link.upload(self._sim)
        self._state = HSFMState((poses.copy(), vels.copy()), self._tracker.state)
        self._collision_radius = None     # Simulation sets it to its robot model's radius (collision list of the fused step)
        self._fused = None
        self._det_cfg = None              # (max_dist, fov, return code) of the detector the fused step evaluates, if any

    @property
    def state(self) -> HSFMState:
        return self._state

    def configure_fused_detector(self, max_dist: float, fov: float, code: int):
        """Simulation hands over the configuration of its PedestrianDetector: the step kernel's epilogue then evaluates the
        reading for the robot pose of every step, and the sensor picks it up from the state object (``_fused_sense``)."""
        self._det_cfg = (float(max_dist), float(fov), int(code))
        self._sim.set_detector(self._det_cfg[0], self._det_cfg[1], self._det_cfg[2], enabled=True)

    def fused_collisions(self, robot_pose: np.ndarray, radius: float):
        """Robot-pedestrian collision list (core/_simulation.py:141-143) of the step just taken, if the kernel's fused
        epilogue evaluated it for exactly this robot pose and radius; else None (the caller runs its own pass)."""
        f = self._fused
        if f is None or f[1] != radius or not np.array_equal(f[0], robot_pose):
            return None
        return list(f[2])

    def step(self, dt: float, robot_pose: Optional[np.ndarray], robot_velocity: Optional[np.ndarray]):
        if robot_pose is None:
            assert robot_velocity is None
        elif robot_velocity is None:
            assert robot_pose is None
        if self._noise_std is not None:                        # :288-289, same global RNG draw
            self._sim.set_accel_noise(np.random.normal(0, self._noise_std, (self._n, 2))[None])
        # one call, one synchronisation: robot upload, step, state + indices + collision list + flags back.  The robot is
        # handed over even when the pedestrians do not see it (_hsfm.py:279-284: visible = False keeps it out of the
        # forces) so that the fused epilogue can evaluate the collision list Simulation asks for next.
        radius = self._collision_radius if self._collision_radius is not None else self._robot_radius
        if robot_pose is not None:
            rp = np.asarray(robot_pose, dtype=np.float64)
            tick = self._sim.step_tick(dt, 1, rp[None], np.asarray(robot_velocity, dtype=np.float64)[None], radius,
                                       self._robot_visible)
        else:
            tick = self._sim.step_tick(dt, 1, None)
        if tick["nonfinite"][0]:   # the reference's numba inverse raises on non-finite headings
            raise np.linalg.LinAlgError("Array must not contain infs or NaNs")
        self._link.after_step(self._sim, tick)
        # the state object owns copies of the two (N, 3) arrays of the tick block; per-pedestrian dicts are built on access
        self._state = HSFMState((tick["poses"][0].copy(), tick["vels"][0].copy()), self._tracker.state)
        if robot_pose is not None:
            self._fused = (rp.copy(), radius, _mask_ids(tick["coll_mask"][0], self._n))
            if self._det_cfg is not None:
                self._state._fused_sense = (self._fused[0], self._det_cfg, _mask_ids(tick["det_mask"][0], self._n),
                                            tick["det_vals"][0].copy())
        else:
            self._fused = None

    def reset_to_state(self, state: HSFMState):
        self._state = state
        self._fused = None
        self._tracker.reset_to_state(state.waypoints)
        poses, vels = state.poses, state.velocities
        self._sim.set_state(np.stack([poses[i] for i in range(self._n)])[None], np.stack([vels[i] for i in range(self._n)])[None])
        self._sim.reset_stats()
        self._link.upload(self._sim)


# ---------------------------------------------------------------------------------------------------------
# ORCA
# ---------------------------------------------------------------------------------------------------------
class ORCAState(_ArrayBackedState):
    pass


class ORCAPedestriansModel(AbstractPedestriansModel):
    """GPU-backed ORCA with the constructor of pedestrians/_orca.py:30-39.  dt, robot_pose and
    robot_velocity passed to step() are ignored, like in the reference (:95-98)."""

    def __init__(self, dt: float, waypoint_tracker: AbstractWaypointTracker, n_pedestrians: int, params=ORCAParams(),
                 initial_poses: Optional[np.ndarray] = None, initial_velocities: Optional[np.ndarray] = None,
                 preferred_speeds: Optional[np.ndarray] = None, max_speeds: Optional[np.ndarray] = None,
                 robot_visible: bool = False, device: int = 0):
        poses, vels = _initial_arrays(waypoint_tracker, n_pedestrians, initial_poses, initial_velocities)
        if max_speeds is None:
            max_speeds = np.array([params.default_max_speed for _ in range(n_pedestrians)])
        else:
            assert max_speeds.shape == (n_pedestrians,), f"max_speeds must have shape (n_pedestrians,), got shape {max_speeds.shape}"
        assert (np.linalg.norm(vels[:, :2], axis=1) <= max_speeds).all(), "Initial speeds must be <= all corresponding max_speeds"
        if preferred_speeds is not None:
            assert preferred_speeds.shape == (n_pedestrians,), f"preferred_speeds must have shape (n_pedestrians,), got shape {preferred_speeds.shape}"
            assert (preferred_speeds <= max_speeds).all(), "All preferred_speeds must be <= corresponding max_speeds"
        else:
            preferred_speeds = max_speeds.copy()
        self._n = poses.shape[0]
        self._ids = list(range(self._n))
        self._tracker = waypoint_tracker
        if self._tracker.state is None:
            self._tracker.resample_all({i: poses[i] for i in range(self._n)})
        self._robot_visible = robot_visible
        self._sim = BatchedOrca(dt, 1, self._n, params, device=device)
        self._sim.set_agents(poses[None, :, :2], vels[None, :, :2], max_speeds[None], preferred_speeds[None])
        # FixedWaypointTracker: the hand-off runs on the device (table upload); any other tracker (the reference accepts every
        # AbstractWaypointTracker, _orca.py:30-39,106-107) runs on the host and hands its current waypoints over every step
        self._device_tracker = isinstance(waypoint_tracker, FixedWaypointTracker)
        self._upload_waypoints()
        self._state = ORCAState((poses.copy(), vels.copy()), self._tracker.state)

    def _upload_waypoints(self):
        if self._device_tracker:
            ids, table = self._tracker.table()
            idx = np.array([self._tracker.state.current_indices[i] for i in range(self._n)], dtype=np.int32)[None]
            self._sim.set_waypoints_fixed(table[None], start_index=idx, reach_distance=self._tracker.reach_distance, loop=self._tracker.loop)
        else:
            cur = self._tracker.state.current_waypoints
            self._sim.set_waypoints_current(np.stack([np.asarray(cur[i], dtype=np.float64)[:2] for i in range(self._n)])[None])

    @property
    def state(self) -> ORCAState:
        return self._state

    def step(self, dt: float = None, robot_pose: Optional[np.ndarray] = None, robot_velocity: Optional[np.ndarray] = None):
        poses, vels = self._sim.tick(1)                # doStep + the agents back: one call, one synchronisation
        if self._device_tracker and self._tracker._stacked is not None and list(self._tracker._stacked[0]) == self._ids:
            # host mirror of the device hand-off on the position array; the tracker state object is built on demand
            self._tracker.update_from_array(poses[0, :, :2])
            self._state = ORCAState((poses[0], vels[0]), self._tracker.state_later())
            return
        if self._tracker.state is not None:
            # FixedWaypointTracker: host mirror of the device hand-off; any other tracker: THE hand-off (_orca.py:106-107),
            # followed by _update_preferred_velocities (:118-127) on the device from the new current waypoints
            self._tracker.update_waypoints({i: poses[0, i] for i in range(self._n)})
            if not self._device_tracker:
                self._upload_waypoints()
        self._state = ORCAState((poses[0], vels[0]), self._tracker.state)      # get_agents returned fresh arrays

    def reset_to_state(self, state: ORCAState):
        # like the reference, the rvo2 agents themselves are not reset (_orca.py:114-116); the tracker is: the next doStep
        # still runs on the old preferred velocities, the hand-off and _update_preferred_velocities (:106-112,118-127) that
        # follow it see the reset tracker -- so the device copy of the tracker (indices, current waypoints) is re-uploaded
        # with the preferred velocities left as they are
        self._state = state
        self._tracker.reset_to_state(state.waypoints)
        self._sim.set_pref_refresh(False)
        self._upload_waypoints()
        self._sim.set_pref_refresh(True)


# ---------------------------------------------------------------------------------------------------------
# ReplayPedestriansPolicy (pedestrians/_replay.py): recorded trajectories, one frame per step.  No dynamics: the
# frames are handed to Simulation, whose collision list and sensors run on the device like for every other model.
# ---------------------------------------------------------------------------------------------------------
class ReplayPedestriansState(AbstractPedestriansModelState):
    def __init__(self, step: int, pedestrians: Dict[int, Tuple[np.ndarray, np.ndarray]]):
        super().__init__(pedestrians, None)
        assert isinstance(step, int) and step >= 0, f"step must be int >= 0, got {step}"        # _replay.py:10
        self._current_step = step

    @property
    def current_step(self) -> int:
        return self._current_step


class ReplayPedestriansPolicy(AbstractPedestriansModel):
    def __init__(self, poses: np.ndarray, velocities: np.ndarray, loop: bool = False) -> None:
        assert len(poses.shape) == 3 and poses.shape[2] == 3, \
            f"Pedestrians poses must have shape (T, n_pedestrians, 3), got {poses.shape}"
        assert len(velocities.shape) == 3 and velocities.shape[2] == 3, \
            f"Pedestrians velocities must have shape (T, n_pedestrians, 3), got {velocities.shape}"
        assert poses.shape[0] == velocities.shape[0] and poses.shape[1] == velocities.shape[1], \
            (f"Poses and velocities must match in time and pedestrians dimension, "
             f"got shapes {poses.shape} and {velocities.shape}")
        self._poses, self._velocities, self._loop = poses.copy(), velocities.copy(), loop
        self._max_step, self._n = poses.shape[0] - 1, poses.shape[1]
        self._current_step = 0
        self._state = self._frame(0)
        # Under this package's Simulation (which sets _collision_radius and the detector configuration) the recorded frames
        # live on the device: a tick uploads the robot pose only and one launch returns the collision list and the detector
        # reading of the frame (pms_replay_tick) instead of a pose upload + sense pass per tick.
        self._collision_radius = None
        self._det_cfg = None
        self._sim = None
        self._fused = None

    def _frame(self, step: int) -> ReplayPedestriansState:
        return ReplayPedestriansState(step, {i: (self._poses[step, i], self._velocities[step, i]) for i in range(self._n)})

    @property
    def state(self) -> ReplayPedestriansState:
        return self._state

    def configure_fused_detector(self, max_dist: float, fov: float, code: int):
        self._det_cfg = (float(max_dist), float(fov), int(code))
        if self._sim is not None:
            self._sim.set_detector(*self._det_cfg, enabled=True)

    def fused_collisions(self, robot_pose: np.ndarray, radius: float):
        f = self._fused
        if f is None or f[1] != radius or not np.array_equal(f[0], robot_pose):
            return None
        return list(f[2])

    def _device(self):
        if self._sim is None:
            self._sim = BatchedSimulation(1, self._n, precision="fp64")
            self._sim.set_replay(self._poses, self._velocities, self._loop)
            self._sim.replay_frame(self._current_step)
            if self._det_cfg is not None:
                self._sim.set_detector(*self._det_cfg, enabled=True)
        return self._sim

    def step(self, dt: float = None, robot_pose: Optional[np.ndarray] = None, robot_velocity: Optional[np.ndarray] = None):
        self._fused = None
        fused = robot_pose is not None and self._collision_radius is not None      # running under this package's Simulation
        if fused:
            rp = np.asarray(robot_pose, dtype=np.float64)
            coll, det, vals, frame = self._device().replay_tick(rp[None], self._collision_radius)
        if self._current_step == self._max_step:        # _replay.py:46-52: stay on the last frame unless looping
            if not self._loop:
                if fused:
                    self._attach(rp, coll, det, vals, frame)
                return
            self._current_step = 0
        else:
            self._current_step += 1
        self._state = self._frame(self._current_step)
        if fused:
            self._attach(rp, coll, det, vals, frame)

    def _attach(self, rp, coll, det, vals, frame):
        assert frame == self._current_step, (frame, self._current_step)            # device and host follow the same rule
        self._fused = (rp.copy(), self._collision_radius, _mask_ids(coll[0], self._n))
        if self._det_cfg is not None:
            self._state._fused_sense = (self._fused[0], self._det_cfg, _mask_ids(det[0], self._n), vals[0].copy())

    def reset_to_state(self, state: ReplayPedestriansState):
        self._state = state
        self._current_step = state.current_step
        self._fused = None
        if self._sim is not None:
            self._sim.replay_frame(self._current_step)


# ---------------------------------------------------------------------------------------------------------
# ETHPedestriansRecord (pedestrians/_eth.py): the BIWI "eth" / "hotel" recordings (2.5 Hz) played back at the simulation's
# rate -- positions interpolated between two recorded frames, pedestrians that leave before the next frame extrapolated
# with their recorded velocity.  The set of pedestrians changes from frame to frame (ids are the recording's).
# The recordings are data of the reference's distribution (pedestrians/assets/obsmat_<sequence>.txt), not part of this
# package: they are read from `obsmat_path`, $PMS_ETH_ASSETS, or an installed pyminisim's assets directory.
# Under this package's Simulation the interpolated frames live in HBM (slots = the pedestrians present, padding parked
# far away) and a tick is one pms_replay_tick, like ReplayPedestriansPolicy.
# ---------------------------------------------------------------------------------------------------------
class ETHState(AbstractPedestriansModelState):
    def __init__(self, pedestrians: Dict[int, Tuple[np.ndarray, np.ndarray]], current_frame: int, current_sub_step: int):
        super().__init__(pedestrians, None)
        self._current_frame, self._current_sub_step = current_frame, current_sub_step

    @property
    def current_frame(self) -> int:
        return self._current_frame

    @property
    def current_sub_step(self) -> int:
        return self._current_sub_step


def _find_obsmat(sequence: str, obsmat_path: Optional[str]) -> str:
    import importlib.util
    import os
    name = f"obsmat_{sequence}.txt"
    candidates = []
    if obsmat_path is not None:
        candidates.append(os.path.join(obsmat_path, name) if os.path.isdir(obsmat_path) else obsmat_path)
    if os.environ.get("PMS_ETH_ASSETS"):
        candidates.append(os.path.join(os.environ["PMS_ETH_ASSETS"], name))
    try:
        spec = importlib.util.find_spec("pyminisim")
        for root in (spec.submodule_search_locations or []) if spec is not None else []:
            candidates.append(os.path.join(root, "pedestrians", "assets", name))
    except (ImportError, ValueError):
        pass
    for c in candidates:
        if os.path.isfile(c):
            return c
    raise FileNotFoundError(f"{name} not found (looked at {candidates or 'nothing'}): pass obsmat_path= or set PMS_ETH_ASSETS "
                            f"to the directory that holds the reference's pedestrians/assets")


class ETHPedestriansRecord(AbstractPedestriansModel):
    SEQUENCE_ETH = "eth"
    SEQUENCE_HOTEL = "hotel"

    _RECORD_DT = 0.4
    _CENTER_ETH = (3.86457727, 3.82923192)
    _CENTER_HOTEL = (-0.21777709, -3.30031001)
    _PARKED = 1.0e18                  # device slots without a pedestrian: out of every sensor's and the collision test's reach

    def __init__(self, sequence: str, dt: float, start_frame: int = 0, obsmat_path: Optional[str] = None):
        centers = {self.SEQUENCE_ETH: self._CENTER_ETH, self.SEQUENCE_HOTEL: self._CENTER_HOTEL}
        if sequence not in centers:                                                              # _eth.py:38-45
            raise ValueError(f"Unknown sequence name {sequence}, available options: "
                             f"{self.SEQUENCE_ETH} and {self.SEQUENCE_HOTEL}")
        cx, cy = centers[sequence]
        # obsmat columns: frame, id, x, z, y, v_x, v_z, v_y; the scene is mirrored in x about its centre (_eth.py:56-69)
        frames: List[Dict[int, np.ndarray]] = []
        last_frame, seen = None, set()
        with open(_find_obsmat(sequence, obsmat_path)) as f:
            for line in f:
                e = line.split()
                if not e:
                    continue
                frame, ped = int(float(e[0])), int(float(e[1]))
                x, y, vx, vy = -(float(e[2]) - cx), float(e[4]) - cy, -float(e[5]), float(e[7])
                if frame != last_frame:
                    frames.append({})
                    last_frame = frame
                frames[-1][ped] = np.array([x, y, np.arctan2(vy, vx), vx, vy, 0.])
                seen.add(ped)
        assert start_frame < len(frames), \
            f"start_frame must be less than number of frames in sequence which is {len(frames)}"
        self._dt = dt
        self._frames = frames
        self._n_peds = len(seen)
        self._n_frames = len(frames) - 1                        # the last recorded frame only ever serves as "next" (_eth.py:88)
        self._steps_per_gap = int(self._RECORD_DT // dt)
        self._blocks: Dict[int, tuple] = {}
        self._state = ETHState(self._pedestrians_at(0, 0), start_frame, 0)       # frame 0 shown, counter at start_frame (:91)
        self._collision_radius = None
        self._det_cfg = None
        self._sim = None
        self._fused = None

    @property
    def state(self) -> ETHState:
        return self._state

    # ---- the interpolation (_eth.py:113-137), one recorded frame = steps_per_gap simulation frames -----------------
    def _block(self, frame: int):
        """(ids, poses (steps_per_gap, P, 3), velocities (P, 3)) of the simulation frames between recorded frames
        `frame` and `frame + 1`."""
        blk = self._blocks.get(frame)
        if blk is None:
            cur, nxt, n = self._frames[frame], self._frames[frame + 1], self._steps_per_gap
            ids = list(cur.keys())
            poses, vels = np.empty((n, len(ids), 3)), np.empty((len(ids), 3))
            for k, ped in enumerate(ids):
                a = cur[ped]
                if ped in nxt:
                    b = nxt[ped]
                    poses[:, k, :2] = np.linspace(a[:2], b[:2], n)
                    poses[:, k, 2] = angle_linspace(a[2], b[2], n)[:n]
                    vels[k] = (a[3], a[4], wrap_angle(b[2] - a[2]) / self._RECORD_DT)
                else:                                            # leaves before the next recorded frame: constant velocity
                    for s in range(n):
                        poses[s, k, :2] = a[:2] + a[3:5] * (s * self._dt)
                    poses[:, k, 2] = a[2]
                    vels[k] = (a[3], a[4], 0.)
            blk = self._blocks[frame] = (ids, poses, vels)
        return blk

    def _pedestrians_at(self, frame: int, sub_step: int) -> Dict[int, Tuple[np.ndarray, np.ndarray]]:
        ids, poses, vels = self._block(frame)
        return {ped: (poses[sub_step, k].copy(), vels[k].copy()) for k, ped in enumerate(ids)}

    # ---- device playback ---------------------------------------------------------------------------------------------
    def configure_fused_detector(self, max_dist: float, fov: float, code: int):
        self._det_cfg = (float(max_dist), float(fov), int(code))
        if self._sim is not None:
            self._sim.set_detector(*self._det_cfg, enabled=True)

    def fused_collisions(self, robot_pose: np.ndarray, radius: float):
        f = self._fused
        if f is None or f[1] != radius or not np.array_equal(f[0], robot_pose):
            return None
        return list(f[2])

    def _flat(self, frame: int, sub_step: int) -> int:
        return frame * self._steps_per_gap + sub_step

    def _device(self):
        if self._sim is None:
            n, slots = self._steps_per_gap, max(1, max(len(self._frames[f]) for f in range(self._n_frames)))
            poses = np.zeros((self._n_frames * n, slots, 3))
            poses[:, :, 0] = self._PARKED
            poses[:, :, 1] = self._PARKED + 1.0e15 * np.arange(1, slots + 1)
            vels = np.zeros_like(poses)
            for f in range(self._n_frames):
                ids, p, v = self._block(f)
                poses[f * n:(f + 1) * n, :len(ids)] = p
                vels[f * n:(f + 1) * n, :len(ids)] = v
            self._slots = slots
            self._sim = BatchedSimulation(1, slots, precision="fp64")
            self._sim.set_replay(poses, vels, loop=True)        # the recording wraps to its first frame (_eth.py:101-102)
            self._sim.replay_frame(self._flat(self._state.current_frame, self._state.current_sub_step))
            if self._det_cfg is not None:
                self._sim.set_detector(*self._det_cfg, enabled=True)
        return self._sim

    def step(self, dt: float = None, robot_pose: Optional[np.ndarray] = None, robot_velocity: Optional[np.ndarray] = None):
        self._fused = None
        sub, frame = self._state.current_sub_step + 1, self._state.current_frame                 # _eth.py:97-106
        if sub >= self._steps_per_gap:
            sub, frame = 0, frame + 1
            if frame >= self._n_frames:
                frame = 0
        fused = robot_pose is not None and self._collision_radius is not None      # running under this package's Simulation
        if fused:
            rp = np.asarray(robot_pose, dtype=np.float64)
            coll, det, vals, shown = self._device().replay_tick(rp[None], self._collision_radius)
            assert shown == self._flat(frame, sub), (shown, frame, sub)            # device and host follow the same rule
        self._state = ETHState(self._pedestrians_at(frame, sub), frame, sub)
        if fused:
            ids = self._block(frame)[0]
            self._fused = (rp.copy(), self._collision_radius, [ids[k] for k in _mask_ids(coll[0], self._slots) if k < len(ids)])
            if self._det_cfg is not None:
                seen = [k for k in _mask_ids(det[0], self._slots) if k < len(ids)]
                self._state._fused_sense = (self._fused[0], self._det_cfg, [ids[k] for k in seen],
                                            {ids[k]: vals[0, k].copy() for k in seen})

    def reset_to_state(self, state: ETHState):
        self._state = state
        self._fused = None
        if self._sim is not None:
            self._sim.replay_frame(self._flat(state.current_frame, state.current_sub_step))
