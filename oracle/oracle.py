"""ctypes front-end of the CPU oracle (``liboracle``) -- TEST INFRASTRUCTURE ONLY.

Used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py``.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpms_oracle.so")
_lib = None

ROBOT_NONE, ROBOT_EXTERNAL, ROBOT_UNICYCLE, ROBOT_BICYCLE = 0, 1, 2, 3
TRACKER_FIXED, TRACKER_QUEUE = 0, 1
MAP_EMPTY, MAP_CIRCLES, MAP_AABB = 0, 1, 2
DET_POLAR, DET_RELATIVE_ROTATED, DET_RELATIVE, DET_ABSOLUTE = 0, 1, 2, 3

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_uint32)
_lp = C.POINTER(C.c_int64)


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc/g++, OpenMP)."""
    srcs = [os.path.join(_HERE, f) for f in ("pms_oracle.c", "pms_oracle.h", "orca_oracle.cpp", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-s", "-B", "libpms_oracle.so"], check=True)
    return _LIB_PATH


class _Batch(C.Structure):
    _fields_ = [
        ("n_worlds", C.c_int), ("n_peds", C.c_int),
        ("tau", C.c_double), ("A", C.c_double), ("B", C.c_double), ("k1", C.c_double), ("k2", C.c_double),
        ("ko", C.c_double), ("kd", C.c_double), ("k_lambda", C.c_double), ("alpha", C.c_double),
        ("mass", C.c_double), ("ped_radius", C.c_double), ("hsfm_robot_radius", C.c_double),
        ("poses", _dp), ("vels", _dp), ("vmag", _dp),
        ("tracker_kind", C.c_int), ("reach", C.c_double), ("loop", C.c_int), ("n_table", C.c_int),
        ("table", _dp), ("wp_index", _ip), ("cur_wp", _dp),
        ("q_cap", C.c_int), ("queue", _dp), ("q_head", _ip), ("q_count", _ip),
        ("ev_cap", C.c_int), ("ev_count", _ip), ("events", _ip),
        ("robot_model", C.c_int), ("robot_visible", C.c_int), ("robot_radius", C.c_double),
        ("wheel_base", C.c_double),
        ("robot_pose", _dp), ("robot_vel", _dp), ("robot_ctrl", _dp), ("robot_rear", _dp),
        ("controls", _dp), ("n_controls", C.c_int), ("robot_world_collision", _ip),
        ("map_kind", C.c_int), ("map_count", C.c_int), ("map_world_stride", C.c_int), ("map_data", _dp),
        ("det_enabled", C.c_int), ("det_max_dist", C.c_double), ("det_fov", C.c_double),
        ("det_return_type", C.c_int),
        ("det_mask", _up), ("det_vals", _dp), ("coll_mask", _up), ("det_count", _lp), ("coll_count", _lp),
        ("nonfinite", _ip), ("steps_done", C.c_int64),
    ]


class _OrcaBatch(C.Structure):
    _fields_ = [
        ("n_worlds", C.c_int), ("n_agents", C.c_int),
        ("time_step", C.c_float), ("neighbor_dist", C.c_float), ("max_neighbors", C.c_int),
        ("time_horizon", C.c_float), ("radius", C.c_float),
        ("pos", _fp), ("vel", _fp), ("pref", _fp), ("max_speed", _fp), ("pref_speed", _dp),
        ("goal_threshold", C.c_double),
        ("n_table", C.c_int), ("table", _dp), ("wp_index", _ip), ("cur_wp", _dp),
        ("reach", C.c_double), ("loop", C.c_int), ("heading", _dp), ("n_constraints", _ip),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_pair_force.argtypes = [C.c_double, C.c_double, _dp, _dp, _dp, _dp,
                                        C.c_double, C.c_double, C.c_double, C.c_double, _dp]
        _lib.orc_pair_force.restype = None
        _lib.orc_wrap_angle.argtypes = [C.c_double]
        _lib.orc_wrap_angle.restype = C.c_double
        _lib.orc_rollout.argtypes = [C.POINTER(_Batch), C.c_double, C.c_int, C.c_int]
        _lib.orc_rollout.restype = C.c_int
        _lib.orc_sense.argtypes = [C.POINTER(_Batch), C.c_int]
        _lib.orc_sense.restype = None
        _lib.orc_lidar_scan.argtypes = [C.POINTER(_Batch), C.c_int, C.c_double, C.c_double,
                                        C.POINTER(C.c_int32), _dp, C.c_int]
        _lib.orc_lidar_scan.restype = None
        _lib.orc_omni.argtypes = [C.POINTER(_Batch), _dp]
        _lib.orc_omni.restype = None
        _lib.orc_map_is_collision.argtypes = [C.POINTER(_Batch), C.c_int, C.c_double, C.c_double, C.c_double]
        _lib.orc_map_is_collision.restype = C.c_int
        _lib.orc_orca_rollout.argtypes = [C.POINTER(_OrcaBatch), C.c_int, C.c_int]
        _lib.orc_orca_rollout.restype = C.c_int
        _lib.orc_orca_update_pref.argtypes = [C.POINTER(_OrcaBatch)]
        _lib.orc_orca_update_pref.restype = None
        _lib.orc_orca_solve_batch.argtypes = [C.c_int, C.c_int, _fp, _fp, _fp, _fp, _ip]
        _lib.orc_orca_solve_batch.restype = None
        _lib.orc_orca_halfplane_batch.argtypes = [C.c_int, _fp, _fp, _fp, _fp, C.c_float, C.c_float, C.c_float, _fp]
        _lib.orc_orca_halfplane_batch.restype = None
        _lib.orc_orca_solve_batch_f64.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, _ip]
        _lib.orc_orca_solve_batch_f64.restype = None
        _lib.orc_orca_halfplane_batch_f64.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, _dp]
        _lib.orc_orca_halfplane_batch_f64.restype = None
    return _lib


def _f64(a, shape=None):
    a = np.ascontiguousarray(np.array(a, dtype=np.float64, copy=True))
    if shape is not None:
        a = np.ascontiguousarray(np.broadcast_to(a, shape)).copy()
    return a


def _ptr(a, typ):
    return a.ctypes.data_as(typ) if a is not None else None


HSFM_DEFAULTS = dict(tau=0.2, A=2000.0, B=0.08, k1=1.2e5, k2=2.4e5, ko=1.0, kd=500.0, k_lambda=0.3,
                     alpha=3.0, mass=70.0, ped_radius=0.3, hsfm_robot_radius=0.35)


def pair_force(radius_i, radius_j, r_i, r_j, v_i, v_j, A=2000.0, B=0.08, k1=1.2e5, k2=2.4e5):
    out = np.zeros(2)
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (r_i, r_j, v_i, v_j)]
    lib().orc_pair_force(radius_i, radius_j, *[_ptr(x, _dp) for x in a], A, B, k1, k2, _ptr(out, _dp))
    return out


def wrap_angle(a: float) -> float:
    return lib().orc_wrap_angle(float(a))


class OracleBatch:
    """W independent worlds of N pedestrians, stepped on the CPU in fp64."""

    def __init__(self, poses, vels=None, vmag=1.5, params: Optional[dict] = None):
        poses = _f64(poses)
        if poses.ndim == 2:
            poses = poses[None]
        self.W, self.N = poses.shape[:2]
        W, N = self.W, self.N
        self.words = (N + 31) // 32
        self.poses = poses
        self.vels = _f64(np.zeros((W, N, 3)) if vels is None else vels, (W, N, 3))
        self.vmag = _f64(vmag, (W, N))
        self.params = dict(HSFM_DEFAULTS)
        if params:
            self.params.update(params)
        # tracker: default = one static waypoint per ped = own position (caller overrides)
        self.tracker_kind = TRACKER_FIXED
        self.reach, self.loop = 0.3, 0
        self.table = _f64(self.poses[:, :, None, :2])
        self.wp_index = np.zeros((W, N), dtype=np.int32)
        self.cur_wp = _f64(self.poses[:, :, :2])
        self.q_cap = 0
        self.queue = None
        self.q_head = np.zeros((W, N), dtype=np.int32)
        self.q_count = np.zeros((W, N), dtype=np.int32)
        self.ev_cap = 0
        self.ev_count = np.zeros(1, dtype=np.int32)
        self.events = None
        # robot
        self.robot_model, self.robot_visible = ROBOT_NONE, 1
        self.robot_radius, self.wheel_base = 0.35, 1.0
        self.robot_pose = np.zeros((W, 3))
        self.robot_vel = np.zeros((W, 3))
        self.robot_ctrl = np.zeros((W, 2))
        self.robot_rear = np.zeros((W, 3))
        self.controls = None
        self.robot_world_collision = np.zeros(W, dtype=np.int32)
        # map
        self.map_kind, self.map_data, self.map_shared = MAP_EMPTY, None, True
        # sensors
        self.det_enabled, self.det_max_dist, self.det_fov, self.det_return_type = 0, 3.0, np.deg2rad(30.0), DET_POLAR
        self.det_mask = np.zeros((W, self.words), dtype=np.uint32)
        self.det_vals = np.zeros((W, N, 2))
        self.coll_mask = np.zeros((W, self.words), dtype=np.uint32)
        self.det_count = np.zeros(W, dtype=np.int64)
        self.coll_count = np.zeros(W, dtype=np.int64)
        self.nonfinite = np.zeros(W, dtype=np.int32)
        self.steps_done = 0

    # ---- configuration -------------------------------------------------------------------
    def set_waypoints_fixed(self, table, start_index=1, reach=0.3, loop=False):
        """table (W,N,K,2) includes the initial position at row 0 (_fixed_waypoint_tracker.py:45,59)."""
        W, N = self.W, self.N
        self.table = _f64(table, (W, N) + tuple(np.shape(table)[-2:]))
        self.tracker_kind, self.reach, self.loop = TRACKER_FIXED, float(reach), int(bool(loop))
        self.wp_index = np.ascontiguousarray(np.broadcast_to(np.asarray(start_index, dtype=np.int32), (W, N))).copy()
        idx = self.wp_index[..., None, None]
        self.cur_wp = np.ascontiguousarray(np.take_along_axis(self.table, np.broadcast_to(idx, (W, N, 1, 2)), axis=2)[:, :, 0, :])

    def set_waypoint_queue(self, current, queue, reach=0.3, ev_cap=65536):
        W, N = self.W, self.N
        self.tracker_kind, self.reach = TRACKER_QUEUE, float(reach)
        self.cur_wp = _f64(current, (W, N, 2))
        self.queue = _f64(queue)
        self.queue = self.queue.reshape(W, N, -1, 2)
        self.q_cap = self.queue.shape[2]
        self.q_head[:] = 0
        self.q_count[:] = self.q_cap
        self.ev_cap = ev_cap
        self.events = np.zeros((ev_cap, 3), dtype=np.int32)

    def set_robot(self, model, pose, control=None, velocity=None, radius=0.35, visible=True, wheel_base=1.0):
        W = self.W
        self.robot_model, self.robot_visible = int(model), int(bool(visible))
        self.robot_radius, self.wheel_base = float(radius), float(wheel_base)
        self.robot_pose = _f64(pose, (W, 3))
        self.robot_ctrl = _f64(np.zeros(2) if control is None else control, (W, 2))
        if model == ROBOT_EXTERNAL:
            self.robot_vel = _f64(np.zeros(3) if velocity is None else velocity, (W, 3))
        elif model == ROBOT_UNICYCLE:
            th = self.robot_pose[:, 2]
            self.robot_vel = np.stack([self.robot_ctrl[:, 0] * np.cos(th), self.robot_ctrl[:, 0] * np.sin(th),
                                       self.robot_ctrl[:, 1]], axis=1)
        else:
            self.robot_vel = np.zeros((W, 3))
        if model == ROBOT_BICYCLE:  # robot/_bicycle.py:58-62
            th = self.robot_pose[:, 2]
            self.robot_rear = np.stack([self.robot_pose[:, 0] - wheel_base / 2.0 * np.cos(th),
                                        self.robot_pose[:, 1] - wheel_base / 2.0 * np.sin(th), th], axis=1)

    def set_controls(self, controls):
        self.controls = None if controls is None else _f64(controls).reshape(-1, self.W, 2)

    def set_map(self, kind, data=None):
        self.map_kind = int(kind)
        if data is None or kind == MAP_EMPTY:
            self.map_data = None
            return
        data = _f64(data)
        self.map_shared = data.ndim == 2
        self.map_data = data

    def set_detector(self, max_dist=3.0, fov=np.deg2rad(30.0), return_type=DET_POLAR, enabled=True):
        self.det_enabled = int(bool(enabled))
        self.det_max_dist, self.det_fov, self.det_return_type = float(max_dist), float(fov), int(return_type)

    # ---- execution -----------------------------------------------------------------------
    def _struct(self) -> _Batch:
        b = _Batch()
        b.n_worlds, b.n_peds = self.W, self.N
        for k, v in self.params.items():
            setattr(b, k, float(v))
        b.poses, b.vels, b.vmag = _ptr(self.poses, _dp), _ptr(self.vels, _dp), _ptr(self.vmag, _dp)
        b.tracker_kind, b.reach, b.loop = self.tracker_kind, self.reach, self.loop
        b.n_table = self.table.shape[2]
        b.table, b.wp_index, b.cur_wp = _ptr(self.table, _dp), _ptr(self.wp_index, _ip), _ptr(self.cur_wp, _dp)
        b.q_cap, b.queue = self.q_cap, _ptr(self.queue, _dp)
        b.q_head, b.q_count = _ptr(self.q_head, _ip), _ptr(self.q_count, _ip)
        b.ev_cap, b.ev_count, b.events = self.ev_cap, _ptr(self.ev_count, _ip), _ptr(self.events, _ip)
        b.robot_model, b.robot_visible = self.robot_model, self.robot_visible
        b.robot_radius, b.wheel_base = self.robot_radius, self.wheel_base
        b.robot_pose, b.robot_vel = _ptr(self.robot_pose, _dp), _ptr(self.robot_vel, _dp)
        b.robot_ctrl, b.robot_rear = _ptr(self.robot_ctrl, _dp), _ptr(self.robot_rear, _dp)
        b.controls = _ptr(self.controls, _dp)
        b.n_controls = 0 if self.controls is None else self.controls.shape[0]
        b.robot_world_collision = _ptr(self.robot_world_collision, _ip)
        b.map_kind = self.map_kind
        if self.map_data is not None:
            b.map_count = self.map_data.shape[-2]
            b.map_world_stride = 0 if self.map_shared else self.map_data.shape[-2] * self.map_data.shape[-1]
            b.map_data = _ptr(self.map_data, _dp)
        else:
            b.map_count, b.map_world_stride, b.map_data = 0, 0, None
        b.det_enabled, b.det_max_dist, b.det_fov = self.det_enabled, self.det_max_dist, self.det_fov
        b.det_return_type = self.det_return_type
        b.det_mask, b.det_vals, b.coll_mask = _ptr(self.det_mask, _up), _ptr(self.det_vals, _dp), _ptr(self.coll_mask, _up)
        b.det_count, b.coll_count = _ptr(self.det_count, _lp), _ptr(self.coll_count, _lp)
        b.nonfinite, b.steps_done = _ptr(self.nonfinite, _ip), self.steps_done
        return b

    def rollout(self, dt: float, n_steps: int, n_threads: int = 1):
        self.ev_count[0] = 0
        b = self._struct()
        lib().orc_rollout(C.byref(b), float(dt), int(n_steps), int(n_threads))
        self.steps_done = int(b.steps_done)
        return self

    def sense(self):
        b = self._struct()
        lib().orc_sense(C.byref(b), 1)
        return self

    def events_list(self):
        n = min(int(self.ev_count[0]), self.ev_cap)
        return self.events[:n].copy() if self.events is not None else np.zeros((0, 3), dtype=np.int32)

    def lidar(self, n_beams=100, max_dist=8.0, resolution=0.01):
        hit = np.zeros((self.W, n_beams), dtype=np.int32)
        pts = np.zeros((self.W, n_beams, 2))
        b = self._struct()
        lib().orc_lidar_scan(C.byref(b), n_beams, float(max_dist), float(resolution),
                             hit.ctypes.data_as(C.POINTER(C.c_int32)), _ptr(pts, _dp), 1)
        return hit, pts

    def omni(self):
        out = np.zeros(self.W)
        b = self._struct()
        lib().orc_omni(C.byref(b), _ptr(out, _dp))
        return out

    def episode(self, dt: float, n_steps: int, schedule: dict, lidar=(100, 8.0, 0.01), controls=None):
        """n_steps single steps with the reference's sensor scheduler (Simulation._update_sensors_state,
        core/_simulation.py:162-185): per scheduled kind, hold += dt after every step; hold >= period -> reading, hold = 0.
        schedule: {"detector" | "lidar" | "omni" | "collisions": [period, hold_time]} (hold_time is updated in place).
        Returns {kind: {"steps": [...], records...}} with one record per reading, in firing order."""
        out = {k: {"steps": []} for k in schedule}
        ctr = None if controls is None else _f64(controls).reshape(-1, self.W, 2)
        for t in range(n_steps):
            if ctr is not None:
                self.set_controls(ctr[min(t, len(ctr) - 1)][None])
            self.rollout(dt, 1)
            for kind, ph in schedule.items():
                ph[1] = ph[1] + dt
                if not ph[1] >= ph[0]:
                    continue
                ph[1] = 0.0
                rec = out[kind]
                rec["steps"].append(t)
                if kind == "detector":
                    rec.setdefault("mask", []).append(self.det_mask.copy())
                    rec.setdefault("vals", []).append(self.det_vals.copy())
                elif kind == "collisions":
                    rec.setdefault("mask", []).append(self.coll_mask.copy())
                elif kind == "lidar":
                    hit, pts = self.lidar(*lidar)
                    rec.setdefault("hit", []).append(hit)
                    rec.setdefault("points", []).append(pts)
                elif kind == "omni":
                    rec.setdefault("dist", []).append(self.omni())
        if ctr is not None:
            self.set_controls(None)
        return {k: {n: np.asarray(v) for n, v in rec.items()} for k, rec in out.items()}

    def map_is_collision(self, w, x, y, radius):
        b = self._struct()
        return bool(lib().orc_map_is_collision(C.byref(b), int(w), float(x), float(y), float(radius)))

    # helpers mirroring the reference's dict outputs
    def detections(self, w=0):
        out = {}
        for i in range(self.N):
            if (self.det_mask[w, i >> 5] >> (i & 31)) & 1:
                out[i] = (float(self.det_vals[w, i, 0]), float(self.det_vals[w, i, 1]))
        return out

    def collisions(self, w=0):
        return [i for i in range(self.N) if (self.coll_mask[w, i >> 5] >> (i & 31)) & 1]


class OracleOrca:
    """W worlds of N ORCA agents (fp32 agent state like rvo2), FixedWaypointTracker semantics."""

    def __init__(self, dt, poses, table, vels=None, pref_speed=None, max_speed=None, loop=False,
                 neighbor_dist=1.0, max_neighbors=10, time_horizon=2.0, radius=0.3, default_max_speed=1.7,
                 goal_threshold=0.2, reach=0.3, honor_agent_kwargs=True):
        poses = _f64(poses)
        if poses.ndim == 2:
            poses = poses[None]
        self.W, self.N = W, N = poses.shape[:2]
        self.dt, self.neighbor_dist, self.max_neighbors = float(dt), float(neighbor_dist), int(max_neighbors)
        self.time_horizon, self.radius = float(time_horizon), float(radius)
        self.pos = np.ascontiguousarray(poses[..., :2], dtype=np.float32)
        v0 = np.zeros((W, N, 2)) if vels is None else np.asarray(vels, dtype=np.float64).reshape(W, N, -1)[..., :2]
        ms = np.full((W, N), default_max_speed) if max_speed is None else np.broadcast_to(max_speed, (W, N))
        if not honor_agent_kwargs:
            v0 = np.zeros((W, N, 2))
            ms_agent = np.full((W, N), default_max_speed)
        else:
            ms_agent = ms
        self.vel = np.ascontiguousarray(v0, dtype=np.float32)
        self.max_speed = np.ascontiguousarray(ms_agent, dtype=np.float32)
        self.pref_speed = _f64(ms if pref_speed is None else pref_speed, (W, N))
        self.pref = np.zeros((W, N, 2), dtype=np.float32)
        self.goal_threshold, self.reach, self.loop = float(goal_threshold), float(reach), int(bool(loop))
        self.table = _f64(table, (W, N) + tuple(np.shape(table)[-2:]))
        self.wp_index = np.ones((W, N), dtype=np.int32)
        self.cur_wp = np.ascontiguousarray(self.table[:, :, 1, :])
        self.heading = np.ascontiguousarray(poses[..., 2])
        self.n_constraints = np.zeros((W, N), dtype=np.int32)
        lib().orc_orca_update_pref(C.byref(self._struct()))  # _orca.py:86

    def _struct(self):
        b = _OrcaBatch()
        b.n_worlds, b.n_agents = self.W, self.N
        b.time_step, b.neighbor_dist, b.max_neighbors = self.dt, self.neighbor_dist, self.max_neighbors
        b.time_horizon, b.radius = self.time_horizon, self.radius
        b.pos, b.vel, b.pref = _ptr(self.pos, _fp), _ptr(self.vel, _fp), _ptr(self.pref, _fp)
        b.max_speed, b.pref_speed = _ptr(self.max_speed, _fp), _ptr(self.pref_speed, _dp)
        b.goal_threshold = self.goal_threshold
        b.n_table, b.table = self.table.shape[2], _ptr(self.table, _dp)
        b.wp_index, b.cur_wp = _ptr(self.wp_index, _ip), _ptr(self.cur_wp, _dp)
        b.reach, b.loop, b.heading = self.reach, self.loop, _ptr(self.heading, _dp)
        b.n_constraints = _ptr(self.n_constraints, _ip)
        return b

    def rollout(self, n_steps, n_threads=1):
        lib().orc_orca_rollout(C.byref(self._struct()), int(n_steps), int(n_threads))
        return self


def orca_solve_batch(lines, max_speed, pref, fp64=False):
    """Test hook: the oracle's 2-D program + relaxed fallback for m problems of n half-planes {point, dir} each.  fp32 is the
    oracle; fp64=True runs the fp64 instantiation of the same code (tests only).
    Returns (v (m,2), fail (m,) int32: index of the half-plane the 2-D program failed at, n = feasible)."""
    dt, pt = (np.float64, _dp) if fp64 else (np.float32, _fp)
    lines = np.ascontiguousarray(lines, dtype=dt)
    m, n = lines.shape[:2]
    ms = np.ascontiguousarray(np.broadcast_to(max_speed, (m,)), dtype=dt)
    pr = np.ascontiguousarray(pref, dtype=dt).reshape(m, 2)
    v = np.zeros((m, 2), dtype=dt)
    fail = np.zeros(m, dtype=np.int32)
    f = lib().orc_orca_solve_batch_f64 if fp64 else lib().orc_orca_solve_batch
    f(m, n, _ptr(lines, pt), _ptr(ms, pt), _ptr(pr, pt), _ptr(v, pt), _ptr(fail, _ip))
    return v, fail


def orca_halfplane_batch(pa, po, va, vo, radius=0.3, time_horizon=2.0, time_step=0.01, fp64=False):
    """Test hook: the ORCA half-plane {point, dir} agent a (pa, va) takes on for neighbour o (po, vo), m pairs."""
    dt, pt = (np.float64, _dp) if fp64 else (np.float32, _fp)
    arrs = [np.ascontiguousarray(x, dtype=dt).reshape(-1, 2) for x in (pa, po, va, vo)]
    m = arrs[0].shape[0]
    out = np.zeros((m, 4), dtype=dt)
    f = lib().orc_orca_halfplane_batch_f64 if fp64 else lib().orc_orca_halfplane_batch
    f(m, *[_ptr(x, pt) for x in arrs], float(radius), float(time_horizon), float(time_step), _ptr(out, pt))
    return out
