"""Batched world API over libpms_b200: W independent worlds of N pedestrians on one GPU.

This is new surface (the reference has no batching concept); the single-world
drop-in classes in ``pyminisim_b200.pedestrians`` / ``.sensors`` / ``.core`` are thin views over it.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _capi as capi
from ._capi import check, dptr, f64, fptr, iptr, lptr, uptr


@dataclass
class HSFMParams:
    """Same fields and defaults as pyminisim.pedestrians.HSFMParams (_hsfm.py:20-42)."""
    tau: float = 0.2
    A: float = 2000
    B: float = 0.08
    k_1: float = 1.2 * 10 ** 5
    k_2: float = 2.4 * 10 ** 5
    k_o: float = 1.
    k_d: float = 500.
    k_lambda: float = 0.3
    alpha: float = 3.

    @staticmethod
    def create_default() -> "HSFMParams":
        return HSFMParams()


class BatchedSimulation:
    """W worlds x N pedestrians, HSFM dynamics + robot + detector/collision pass fused on the GPU."""

    def __init__(self, n_worlds: int, n_peds: int, precision: str = "fp32", device: int = 0,
                 hsfm_params: Optional[HSFMParams] = None, pedestrian_mass: float = 70.0,
                 pedestrian_radius: float = 0.3, hsfm_robot_radius: float = 0.35):
        self._lib = capi.load()
        self.W, self.N = int(n_worlds), int(n_peds)
        self.words = (self.N + 31) // 32
        self.precision = precision
        self.device = int(device)
        p = hsfm_params or HSFMParams()
        cp = capi.HsfmParamsC(float(p.tau), float(p.A), float(p.B), float(p.k_1), float(p.k_2), float(p.k_o),
                              float(p.k_d), float(p.k_lambda), float(p.alpha), float(pedestrian_mass),
                              float(pedestrian_radius), float(hsfm_robot_radius))
        self._async_io, self._pending_nf = False, []
        self._h = C.c_void_p()
        check(self._lib.pms_create(int(device), self.W, self.N, capi.PRECISIONS[precision], C.byref(cp), C.byref(self._h)))
        self.robot_model = capi.ROBOT_NONE

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.pms_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state ----------------------------------------------------------------------------------
    def set_state(self, poses, velocities=None):
        """Host arrays (W, N, 3).  float32 input takes the float32 interface (half the bytes over PCIe), anything else
        the reference's float64."""
        if getattr(poses, "dtype", None) == np.float32 and (velocities is None or getattr(velocities, "dtype", None) == np.float32):
            p32 = np.ascontiguousarray(poses, dtype=np.float32).reshape(self.W, self.N, 3)
            v32 = None if velocities is None else np.ascontiguousarray(velocities, dtype=np.float32).reshape(self.W, self.N, 3)
            check(self._lib.pms_set_ped_state_f32(self._h, fptr(p32), fptr(v32)))
            return
        poses = f64(poses, (self.W, self.N, 3))
        vels = None if velocities is None else f64(velocities, (self.W, self.N, 3))
        check(self._lib.pms_set_ped_state(self._h, dptr(poses), dptr(vels)))

    def get_state(self, out_poses=None, out_vels=None, dtype=np.float64) -> Tuple[np.ndarray, np.ndarray]:
        """(poses, velocities) as host arrays (W, N, 3) of ``dtype`` (float64 like the reference, or float32); output arrays
        passed in decide the dtype themselves."""
        if out_poses is not None:
            dtype = out_poses.dtype
        elif out_vels is not None:
            dtype = out_vels.dtype
        poses = np.empty((self.W, self.N, 3), dtype=dtype) if out_poses is None else out_poses
        vels = np.empty((self.W, self.N, 3), dtype=dtype) if out_vels is None else out_vels
        if np.dtype(dtype) == np.float32:
            check(self._lib.pms_get_ped_state_f32(self._h, fptr(poses), fptr(vels)))
        else:
            check(self._lib.pms_get_ped_state(self._h, dptr(poses), dptr(vels)))
        return poses, vels

    # ---- device-resident I/O (callers that stay on the GPU: torch tensors, CuPy arrays, raw pointers) --------------------
    def set_state_device(self, poses_ptr: int, vels_ptr: Optional[int], f32: bool):
        """``poses_ptr`` / ``vels_ptr``: device addresses of (W, N, 3) float32 / float64 arrays on this handle's device."""
        check(self._lib.pms_set_ped_state_device(self._h, C.c_void_p(poses_ptr), C.c_void_p(vels_ptr or 0) if vels_ptr else None, int(bool(f32))))

    def get_state_device(self, poses_ptr: Optional[int], vels_ptr: Optional[int], f32: bool):
        check(self._lib.pms_get_ped_state_device(self._h, C.c_void_p(poses_ptr) if poses_ptr else None,
                                                 C.c_void_p(vels_ptr) if vels_ptr else None, int(bool(f32))))

    def set_state_torch(self, poses, velocities=None):
        """State from CUDA torch tensors (W, N, 3), float32 or float64, no host round trip.  The conversion kernel runs on
        the handle's stream: the tensors must be complete (synchronise or order the producer stream first)."""
        import torch
        assert poses.is_cuda and poses.is_contiguous() and poses.shape == (self.W, self.N, 3) and poses.dtype in (torch.float32, torch.float64)
        if velocities is not None:
            assert velocities.is_cuda and velocities.is_contiguous() and velocities.shape == poses.shape and velocities.dtype == poses.dtype
        self.set_state_device(poses.data_ptr(), None if velocities is None else velocities.data_ptr(), poses.dtype == torch.float32)

    def get_state_torch(self, dtype=None, out=None):
        """(poses, velocities) as CUDA torch tensors (W, N, 3); written on the handle's stream -- call ``sync()`` (or order
        your stream after ``stream``) before reading them."""
        import torch
        if out is None:
            dtype = dtype or torch.float32
            dev = torch.device("cuda", self.device)
            out = (torch.empty((self.W, self.N, 3), dtype=dtype, device=dev), torch.empty((self.W, self.N, 3), dtype=dtype, device=dev))
        poses, vels = out
        self.get_state_device(poses.data_ptr(), vels.data_ptr(), poses.dtype == torch.float32)
        return poses, vels

    def device_buffer(self, which: int) -> Tuple[int, int]:
        """(device address, bytes) of one of the handle's result buffers (``capi.BUF_*``), zero-copy."""
        p, b = C.c_void_p(), C.c_uint64(0)
        check(self._lib.pms_device_buffer(self._h, int(which), C.byref(p), C.byref(b)))
        return int(p.value or 0), int(b.value)

    def device_tensor(self, which: int, dtype, shape=None):
        """A torch tensor that ALIASES a result buffer (``capi.BUF_*``): no copy, valid as documented for pms_device_buffer."""
        import torch
        ptr, nbytes = self.device_buffer(which)
        if not ptr or not nbytes:
            return None
        np_dt = np.dtype(dtype)

        class _Alias:
            __cuda_array_interface__ = {"shape": (nbytes // np_dt.itemsize,), "typestr": np_dt.str, "data": (ptr, False), "version": 2}
        t = torch.as_tensor(_Alias(), device=torch.device("cuda", self.device))
        return t if shape is None else t.view(*shape)

    def set_speeds(self, vmag):
        check(self._lib.pms_set_ped_speeds(self._h, dptr(f64(vmag, (self.W, self.N)))))

    def set_async_io(self, enabled: bool = True):
        """Transfers (set_state / get_state / get_detector / get_collisions / get_stats) stop synchronising: they are
        enqueued on the handle's stream and complete at ``sync()``.  Host arrays must be page-locked (``_capi.pinned_array``)
        and must not be touched before ``sync()``.  Lets several handles overlap copies with each other's kernels."""
        check(self._lib.pms_set_async_io(self._h, int(bool(enabled))))
        self._async_io = bool(enabled)

    def snapshot_save(self):
        check(self._lib.pms_snapshot_save(self._h))

    def snapshot_restore(self):
        check(self._lib.pms_snapshot_restore(self._h))

    # ---- trackers -------------------------------------------------------------------------------
    def set_waypoints_fixed(self, table, start_index=None, reach_distance: float = 0.3, loop: bool = False):
        table = np.asarray(table, dtype=np.float64)
        n_rows = table.shape[-2]
        table = f64(table, (self.W, self.N, n_rows, 2))
        si = None
        if start_index is not None:
            si = np.ascontiguousarray(np.broadcast_to(np.asarray(start_index, dtype=np.int32), (self.W, self.N)))
        check(self._lib.pms_set_waypoints_fixed(self._h, dptr(table), n_rows, iptr(si), float(reach_distance), int(bool(loop))))

    def set_waypoint_queue(self, current, queue, reach_distance: float = 0.3):
        queue = np.asarray(queue, dtype=np.float64)
        q_len = queue.shape[-2]
        check(self._lib.pms_set_waypoint_queue(self._h, dptr(f64(current, (self.W, self.N, 2))),
                                               dptr(f64(queue, (self.W, self.N, q_len, 2))), q_len, float(reach_distance)))

    def append_waypoints(self, world_ped, points):
        wp = np.ascontiguousarray(world_ped, dtype=np.int32).reshape(-1, 2)
        pts = f64(points).reshape(-1, 2)
        check(self._lib.pms_append_waypoints(self._h, len(wp), iptr(wp), dptr(pts)))

    def pop_waypoint_events(self, cap: int = 1 << 16) -> np.ndarray:
        ev = np.zeros((cap, 3), dtype=np.int32)
        n = C.c_int(0)
        check(self._lib.pms_pop_waypoint_events(self._h, iptr(ev), cap, C.byref(n)))
        ev = ev[:n.value]
        if len(ev):
            ev = ev[np.lexsort((ev[:, 2], ev[:, 1], ev[:, 0]))]  # (step, world, ped) order = reference order
        return ev

    def get_waypoints(self) -> Tuple[np.ndarray, np.ndarray]:
        cur = np.empty((self.W, self.N, 2))
        idx = np.empty((self.W, self.N), dtype=np.int32)
        check(self._lib.pms_get_waypoints(self._h, dptr(cur), iptr(idx)))
        return cur, idx

    # ---- robot / map / sensors --------------------------------------------------------------------
    def set_robot(self, model: int, pose=None, velocity=None, control=None, radius: float = 0.35,
                  wheel_base: float = 1.0, visible: bool = True):
        self.robot_model = int(model)
        pose_a = None if pose is None else f64(pose, (self.W, 3))
        vel_a = None if velocity is None else f64(velocity, (self.W, 3))
        ctl_a = None if control is None else f64(control, (self.W, 2))
        check(self._lib.pms_set_robot(self._h, int(model), dptr(pose_a), dptr(vel_a), dptr(ctl_a), float(radius),
                                      float(wheel_base), int(bool(visible))))

    def set_robot_controls(self, controls):
        if controls is None:
            check(self._lib.pms_set_robot_controls(self._h, None, 0))
            return
        c = f64(controls).reshape(-1, self.W, 2)
        check(self._lib.pms_set_robot_controls(self._h, dptr(c), c.shape[0]))

    def set_robot_controls_torch(self, controls):
        """Controls (n_rows, W, 2) as a CUDA float64 torch tensor, read in place by the launches that follow (row s drives step
        s, the last row is held): a policy's output reaches the robot model without leaving the device.  The tensor must stay
        alive and complete until those launches have run; ``None`` clears."""
        if controls is None:
            check(self._lib.pms_set_robot_controls_device(self._h, None, 0))
            self._controls_ref = None
            return
        import torch
        assert controls.is_cuda and controls.is_contiguous() and controls.dtype == torch.float64 and controls.shape[1:] == (self.W, 2)
        self._controls_ref = controls
        check(self._lib.pms_set_robot_controls_device(self._h, C.c_void_p(controls.data_ptr()), int(controls.shape[0])))

    def get_robot(self):
        pose = np.empty((self.W, 3))
        vel = np.empty((self.W, 3))
        wc = np.empty(self.W, dtype=np.int32)
        check(self._lib.pms_get_robot(self._h, dptr(pose), dptr(vel), iptr(wc)))
        return pose, vel, wc

    def set_map(self, kind: int, data=None):
        if data is None or kind == capi.MAP_EMPTY:
            check(self._lib.pms_set_map(self._h, capi.MAP_EMPTY, None, 0, 0))
            return
        data = f64(data)
        per_world = data.ndim == 3
        check(self._lib.pms_set_map(self._h, int(kind), dptr(data), data.shape[-2], int(per_world)))

    def set_accel_noise(self, noise):
        check(self._lib.pms_set_accel_noise(self._h, None if noise is None else dptr(f64(noise, (self.W, self.N, 2)))))

    def set_accel_noise_std(self, noise_std: float, seed: int = 0):
        """``noise_std`` of HeadedSocialForceModelPolicy (_hsfm.py:208,288-289) drawn on the device, afresh in every step of
        a fused launch (Philox, keyed by seed / pedestrian / absolute step).  Statistical parity with the reference."""
        check(self._lib.pms_set_accel_noise_std(self._h, float(noise_std), int(seed) & 0xFFFFFFFFFFFFFFFF))

    def set_detector_noise(self, misdetection_prob: float = 0.0, distance_mu: float = 0.0, distance_sigma: float = 0.0,
                           angle_mu: float = 0.0, angle_sigma: float = 0.0, seed: int = 0, enabled: bool = True):
        """PedestrianDetectorNoise (sensors/_pedestrian_detector.py:60-72,104-112) applied on the device by ``get_detector``."""
        check(self._lib.pms_set_detector_noise(self._h, int(bool(enabled)), float(misdetection_prob), float(distance_mu),
                                               float(distance_sigma), float(angle_mu), float(angle_sigma),
                                               int(seed) & 0xFFFFFFFFFFFFFFFF))

    def set_lidar_noise(self, point_mean=None, point_cov=None, drop_prob: float = 0.0, seed: int = 0, enabled: bool = True):
        """LidarSensorNoise (sensors/_lidar.py:52-87) applied on the device by ``lidar_scan``."""
        mean = None if point_mean is None else f64(point_mean, (2,))
        cov = None if point_cov is None else f64(point_cov, (2, 2))
        check(self._lib.pms_set_lidar_noise(self._h, int(bool(enabled)), float(drop_prob), dptr(mean), dptr(cov),
                                            int(seed) & 0xFFFFFFFFFFFFFFFF))

    def set_omni_noise(self, misdetection_prob: float = 0.0, distance_mu: float = 0.0, distance_sigma: float = 0.0,
                       seed: int = 0, enabled: bool = True):
        """OmniObstacleDetectorNoise (sensors/_omni_obstacle_detector.py:25-37,64-69) applied on the device by ``omni_scan``."""
        check(self._lib.pms_set_omni_noise(self._h, int(bool(enabled)), float(misdetection_prob), float(distance_mu),
                                           float(distance_sigma), int(seed) & 0xFFFFFFFFFFFFFFFF))

    def set_detector(self, max_dist: float = 3.0, fov: float = np.deg2rad(30.0), return_type: int = capi.DET_POLAR,
                     enabled: bool = True):
        check(self._lib.pms_detector_config(self._h, int(bool(enabled)), float(max_dist), float(fov), int(return_type)))

    # ---- hot path ---------------------------------------------------------------------------------
    def step(self, dt: float = 0.01, n_steps: int = 1):
        check(self._lib.pms_hsfm_step(self._h, float(dt), int(n_steps)))

    def step_async(self, dt: float = 0.01, n_steps: int = 1):
        check(self._lib.pms_hsfm_step_async(self._h, float(dt), int(n_steps)))

    def sync(self):
        check(self._lib.pms_sync(self._h))
        # asynchronous-I/O mode: pms_get_stats left the raw "never" sentinel (INT32_MAX) in the buffers it filled; after the
        # synchronisation they read like the blocking call's (0 = the world stayed finite)
        for nf in self._pending_nf:
            nf[nf == np.iinfo(np.int32).max] = 0
        self._pending_nf.clear()

    def step_tick(self, dt: float, n_steps: int = 1, robot_pose=None, robot_velocity=None, robot_radius: float = 0.35,
                  visible: bool = True) -> dict:
        """One tick in one call and one host synchronisation (``pms_step_tick``): upload the robot the caller supplies
        (``None`` = no robot), step, and fetch state, waypoint indices, collision list, detector reading, non-finite flag and
        the number of logged waypoint events.  The output arrays are reused from tick to tick: copy what you keep."""
        if not hasattr(self, "_tick"):
            self._tick = dict(poses=np.empty((self.W, self.N, 3)), vels=np.empty((self.W, self.N, 3)),
                              wp_index=np.empty((self.W, self.N), dtype=np.int32),
                              coll_mask=np.empty((self.W, self.words), dtype=np.uint32),
                              det_mask=np.empty((self.W, self.words), dtype=np.uint32), det_vals=np.empty((self.W, self.N, 2)),
                              nonfinite=np.empty(self.W, dtype=np.int32), n_events=np.zeros(1, dtype=np.int32))
            t = self._tick
            self._tick_ptrs = (dptr(t["poses"]), dptr(t["vels"]), iptr(t["wp_index"]), uptr(t["coll_mask"]), uptr(t["det_mask"]),
                               dptr(t["det_vals"]), iptr(t["nonfinite"]), iptr(t["n_events"]))
        if robot_pose is None:
            self.robot_model = capi.ROBOT_NONE
            rp = rv = None
        else:
            self.robot_model = capi.ROBOT_EXTERNAL
            rp = f64(robot_pose, (self.W, 3))
            rv = None if robot_velocity is None else f64(robot_velocity, (self.W, 3))
        check(self._lib.pms_step_tick(self._h, float(dt), int(n_steps), dptr(rp), dptr(rv), float(robot_radius), int(bool(visible)),
                                      *self._tick_ptrs))
        return self._tick

    @property
    def stream(self) -> int:
        return int(self._lib.pms_stream(self._h) or 0)

    def sense(self):
        check(self._lib.pms_sense(self._h))

    def sense_tick(self, poses, robot_pose, robot_radius: float = 0.35):
        """Collision list + detector reading for the given poses and robot pose in one call (``pms_sense_tick``)."""
        self.robot_model = capi.ROBOT_EXTERNAL
        coll = np.empty((self.W, self.words), dtype=np.uint32)
        det = np.empty((self.W, self.words), dtype=np.uint32)
        vals = np.empty((self.W, self.N, 2))
        check(self._lib.pms_sense_tick(self._h, dptr(f64(poses, (self.W, self.N, 3))), dptr(f64(robot_pose, (self.W, 3))),
                                       float(robot_radius), uptr(coll), uptr(det), dptr(vals)))
        return coll, det, vals

    def get_detector(self, out_mask=None, out_vals=None) -> Tuple[np.ndarray, np.ndarray]:
        mask = np.empty((self.W, self.words), dtype=np.uint32) if out_mask is None else out_mask
        vals = np.empty((self.W, self.N, 2)) if out_vals is None else out_vals
        check(self._lib.pms_get_detector(self._h, uptr(mask), dptr(vals)))
        return mask, vals

    def get_detector_f32(self, out_mask=None, out_vals=None) -> Tuple[np.ndarray, np.ndarray]:
        """The exact reading with float32 values (half the bytes of ``get_detector``)."""
        mask = np.empty((self.W, self.words), dtype=np.uint32) if out_mask is None else out_mask
        vals = np.empty((self.W, self.N, 2), dtype=np.float32) if out_vals is None else out_vals
        check(self._lib.pms_get_detector_f32(self._h, uptr(mask), fptr(vals)))
        return mask, vals

    def get_collisions(self, out_mask=None) -> np.ndarray:
        mask = np.empty((self.W, self.words), dtype=np.uint32) if out_mask is None else out_mask
        check(self._lib.pms_get_collisions(self._h, uptr(mask)))
        return mask

    def get_stats(self, out=None):
        det, col, nf = out if out is not None else (np.empty(self.W, dtype=np.int64), np.empty(self.W, dtype=np.int64),
                                                    np.empty(self.W, dtype=np.int32))
        check(self._lib.pms_get_stats(self._h, lptr(det), lptr(col), iptr(nf)))
        if self._async_io:
            self._pending_nf.append(nf)
        return det, col, nf

    def reset_stats(self):
        check(self._lib.pms_reset_stats(self._h))

    def lidar_scan(self, n_beams: int = 100, max_dist: float = 8.0, resolution: float = 0.01):
        hit = np.empty((self.W, n_beams), dtype=np.int32)
        pts = np.empty((self.W, n_beams, 2))
        check(self._lib.pms_lidar_scan(self._h, int(n_beams), float(max_dist), float(resolution), iptr(hit), dptr(pts)))
        return hit, pts

    def omni_scan(self) -> np.ndarray:
        out = np.empty(self.W)
        check(self._lib.pms_omni_scan(self._h, dptr(out)))
        return out

    # ---- recorded trajectories (ReplayPedestriansPolicy, pedestrians/_replay.py) as a batched gather -----------------------
    def set_replay(self, poses, velocities, loop: bool = False):
        """Frames (T, W, N, 3) -- or (T, N, 3), shared by every world -- uploaded once; frame 0 becomes the state."""
        poses, velocities = np.asarray(poses, dtype=np.float64), np.asarray(velocities, dtype=np.float64)
        if poses.ndim == 3:
            poses = np.broadcast_to(poses[:, None], (poses.shape[0], self.W, self.N, 3))
            velocities = np.broadcast_to(velocities[:, None], (velocities.shape[0], self.W, self.N, 3))
        T = poses.shape[0]
        check(self._lib.pms_set_replay(self._h, dptr(f64(poses, (T, self.W, self.N, 3))), dptr(f64(velocities, (T, self.W, self.N, 3))),
                                       T, int(bool(loop))))

    def replay_step(self, dt: float = 0.01, n_steps: int = 1):
        """n_steps frames under the robot model and the sensors in one launch (``pms_replay_step``)."""
        check(self._lib.pms_replay_step(self._h, float(dt), int(n_steps)))

    def replay_frame(self, set_frame: int = -1) -> int:
        f = C.c_int(0)
        check(self._lib.pms_replay_frame(self._h, int(set_frame), C.byref(f)))
        return f.value

    def replay_tick(self, robot_pose=None, robot_radius: float = 0.35, n_steps: int = 1):
        """One tick of the drop-in replay policy: (collision mask, detector mask, detector values, frame)."""
        coll = np.empty((self.W, self.words), dtype=np.uint32)
        det = np.empty((self.W, self.words), dtype=np.uint32)
        vals = np.empty((self.W, self.N, 2))
        frame = np.zeros(1, dtype=np.int32)
        rp = None if robot_pose is None else f64(robot_pose, (self.W, 3))
        self.robot_model = capi.ROBOT_NONE if rp is None else capi.ROBOT_EXTERNAL
        check(self._lib.pms_replay_tick(self._h, int(n_steps), dptr(rp), float(robot_radius), uptr(coll), uptr(det), dptr(vals), iptr(frame)))
        return coll, det, vals, int(frame[0])

    # ---- sensors inside a fused launch (the reference's hold-time scheduler on the device) -----------
    def set_lidar(self, n_beams: int = 100, max_dist: float = 8.0, resolution: float = 0.01):
        """LidarSensorConfig (sensors/_lidar.py:8-23) of the lidar readings taken inside fused launches."""
        check(self._lib.pms_lidar_config(self._h, int(n_beams), float(max_dist), float(resolution)))
        self._lidar_beams = int(n_beams)

    def schedule_sensor(self, kind: int, period: float = 0.0, hold_time: float = 0.0, values: bool = False,
                        enabled: bool = True):
        """Schedule a sensor kind (``capi.SENSOR_*``) inside ``step`` with the reference's hold-time rule
        (core/_simulation.py:162-185): after every step ``hold += dt``; ``hold >= period`` -> reading, ``hold = 0``.
        ``period`` = ``AbstractSensor.period`` (1 / frequency, 0 = every step).  The readings of the last launch are the
        *records* returned by ``sensor_fires`` / ``*_records``.  ``values``: detector records keep the (W, N, 2) values."""
        check(self._lib.pms_sensor_schedule(self._h, int(kind), int(bool(enabled)), float(period), float(hold_time),
                                            capi.SENSOR_VALUES if values else 0))

    def sensor_fires(self, kind: int) -> Tuple[np.ndarray, float]:
        """(0-based launch steps after which the kind was read in the last launch, hold_time left in the scheduler)."""
        n, hold = C.c_int(0), C.c_double(0.0)
        check(self._lib.pms_get_sensor_fires(self._h, int(kind), None, 0, C.byref(n), C.byref(hold)))
        steps = np.empty(n.value, dtype=np.int32)
        if n.value:
            check(self._lib.pms_get_sensor_fires(self._h, int(kind), iptr(steps), n.value, C.byref(n), C.byref(hold)))
        return steps, hold.value

    def _n_records(self, kind: int, first: int, count: Optional[int]) -> int:
        if count is not None:
            return int(count)
        n = C.c_int(0)
        check(self._lib.pms_get_sensor_fires(self._h, int(kind), None, 0, C.byref(n), None))
        return n.value - first

    def detector_records(self, first: int = 0, count: Optional[int] = None, values: bool = False):
        count = self._n_records(capi.SENSOR_DETECTOR, first, count)
        mask = np.empty((count, self.W, self.words), dtype=np.uint32)
        vals = np.empty((count, self.W, self.N, 2)) if values else None
        check(self._lib.pms_get_detector_records(self._h, int(first), count, uptr(mask), dptr(vals)))
        return (mask, vals) if values else mask

    def collision_records(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self._n_records(capi.SENSOR_COLLISIONS, first, count)
        mask = np.empty((count, self.W, self.words), dtype=np.uint32)
        check(self._lib.pms_get_collision_records(self._h, int(first), count, uptr(mask)))
        return mask

    def lidar_records(self, first: int = 0, count: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
        count = self._n_records(capi.SENSOR_LIDAR, first, count)
        hit = np.empty((count, self.W, self._lidar_beams), dtype=np.int32)
        pts = np.empty((count, self.W, self._lidar_beams, 2))
        check(self._lib.pms_get_lidar_records(self._h, int(first), count, iptr(hit), dptr(pts)))
        return hit, pts

    def omni_records(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self._n_records(capi.SENSOR_OMNI, first, count)
        out = np.empty((count, self.W))
        check(self._lib.pms_get_omni_records(self._h, int(first), count, dptr(out)))
        return out

    def launch_info(self) -> dict:
        info = capi.LaunchInfoC()
        check(self._lib.pms_get_launch_info(self._h, C.byref(info)))
        return {n: getattr(info, n) for n, _ in capi.LaunchInfoC._fields_}

    def set_tuning(self, j_split: int = 0, worlds_per_block: int = 0, n_chunks: int = 0, cluster_size: int = 0):
        check(self._lib.pms_set_tuning(self._h, int(j_split), int(worlds_per_block), int(n_chunks), int(cluster_size)))

    # ---- helpers mirroring the reference's dict outputs ---------------------------------------------
    @staticmethod
    def mask_to_ids(mask_row: np.ndarray, n: int) -> List[int]:
        return [i for i in range(n) if (int(mask_row[i >> 5]) >> (i & 31)) & 1]


@dataclass
class ORCAParams:
    """Same fields and defaults as pyminisim.pedestrians.ORCAParams (_orca.py:13-21)."""
    neighbor_dist: float = 1.
    max_neighbors: int = 10
    time_horizon: float = 2.
    time_horizon_obst: float = 2.
    radius = 0.3
    default_max_speed: float = 1.7
    goal_reaching_threshold: float = 0.2


class BatchedOrca:
    """W worlds x N ORCA agents (the rvo2 doStep + adapter of _orca.py:95-127) on one GPU."""

    def __init__(self, dt: float, n_worlds: int, n_agents: int, params: Optional[ORCAParams] = None, device: int = 0):
        self._lib = capi.load()
        self.W, self.N = int(n_worlds), int(n_agents)
        p = params or ORCAParams()
        self.params = p
        cp = capi.OrcaParamsC(float(dt), float(p.neighbor_dist), int(p.max_neighbors), float(p.time_horizon),
                              float(p.radius), float(p.goal_reaching_threshold))
        self._h = C.c_void_p()
        check(self._lib.pms_orca_create(int(device), self.W, self.N, C.byref(cp), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.pms_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_agents(self, positions, velocities=None, max_speeds=None, pref_speeds=None):
        pos = f64(np.asarray(positions, dtype=np.float64)[..., :2], (self.W, self.N, 2))
        vel = None if velocities is None else f64(np.asarray(velocities, dtype=np.float64)[..., :2], (self.W, self.N, 2))
        ms = f64(self.params.default_max_speed if max_speeds is None else max_speeds, (self.W, self.N))
        ps = None if pref_speeds is None else f64(pref_speeds, (self.W, self.N))
        check(self._lib.pms_orca_set_agents(self._h, dptr(pos), dptr(vel), dptr(ms), dptr(ps)))

    def set_waypoints_fixed(self, table, start_index=None, reach_distance: float = 0.3, loop: bool = False):
        table = np.asarray(table, dtype=np.float64)
        n_rows = table.shape[-2]
        table = f64(table, (self.W, self.N, n_rows, 2))
        si = None
        if start_index is not None:
            si = np.ascontiguousarray(np.broadcast_to(np.asarray(start_index, dtype=np.int32), (self.W, self.N)))
        check(self._lib.pms_orca_set_waypoints_fixed(self._h, dptr(table), n_rows, iptr(si), float(reach_distance), int(bool(loop))))

    def set_waypoints_current(self, current):
        """Waypoints of a tracker that runs on the host (any AbstractWaypointTracker): hand over its current waypoints after
        every update; the device-side hand-off is off (_orca.py:106-107,118-127)."""
        cur = f64(np.asarray(current, dtype=np.float64)[..., :2], (self.W, self.N, 2))
        check(self._lib.pms_orca_set_waypoints_current(self._h, dptr(cur)))

    def set_pref_refresh(self, enabled: bool):
        """False: the waypoint setters keep the preferred velocities (ORCAPedestriansModel.reset_to_state, _orca.py:114-116)."""
        check(self._lib.pms_orca_set_pref_refresh(self._h, int(bool(enabled))))

    def step(self, n_steps: int = 1):
        check(self._lib.pms_orca_step(self._h, int(n_steps)))

    def step_async(self, n_steps: int = 1):
        check(self._lib.pms_orca_step_async(self._h, int(n_steps)))

    def tick(self, n_steps: int = 1) -> Tuple[np.ndarray, np.ndarray]:
        """n_steps and the agents back in one call and one synchronisation (``pms_orca_tick``): fresh (W, N, 3) arrays."""
        poses = np.empty((self.W, self.N, 3))
        vels = np.empty((self.W, self.N, 3))
        check(self._lib.pms_orca_tick(self._h, int(n_steps), dptr(poses), dptr(vels)))
        return poses, vels

    def sync(self):
        check(self._lib.pms_sync(self._h))

    @property
    def stream(self) -> int:
        return int(self._lib.pms_stream(self._h) or 0)

    def snapshot_save(self):
        check(self._lib.pms_orca_snapshot_save(self._h))

    def snapshot_restore(self):
        check(self._lib.pms_orca_snapshot_restore(self._h))

    def get_agents(self) -> Tuple[np.ndarray, np.ndarray]:
        poses = np.empty((self.W, self.N, 3))
        vels = np.empty((self.W, self.N, 3))
        check(self._lib.pms_orca_get_agents(self._h, dptr(poses), dptr(vels)))
        return poses, vels

    def launch_info(self) -> dict:
        info = capi.LaunchInfoC()
        check(self._lib.pms_get_launch_info(self._h, C.byref(info)))
        return {n: getattr(info, n) for n, _ in capi.LaunchInfoC._fields_}
