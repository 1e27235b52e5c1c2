"""ctypes binding of libpms_b200.so (include/pms.h).  No fallback: a missing library or a missing
CUDA device is an error the caller sees."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PMS_LIB", os.path.join(HERE, "libpms_b200.so"))  # PMS_LIB: experiment builds only

PMS_FP32, PMS_MIXED, PMS_FP64, PMS_FP32_PLAIN = 0, 1, 2, 3
ROBOT_NONE, ROBOT_EXTERNAL, ROBOT_UNICYCLE, ROBOT_BICYCLE = 0, 1, 2, 3
MAP_EMPTY, MAP_CIRCLES, MAP_AABB = 0, 1, 2
DET_POLAR, DET_RELATIVE_ROTATED, DET_RELATIVE, DET_ABSOLUTE = 0, 1, 2, 3
SENSOR_DETECTOR, SENSOR_LIDAR, SENSOR_OMNI, SENSOR_COLLISIONS = 0, 1, 2, 3
SENSOR_VALUES = 1
# pms_buffer: name -> (id, numpy dtype, trailing shape builder)
BUF_DET_MASK, BUF_COLL_MASK, BUF_DET_VALUES, BUF_ROBOT_POSE, BUF_ROBOT_VELOCITY = 0, 1, 2, 3, 4
BUF_DETECTIONS, BUF_COLLISIONS, BUF_FIRST_NONFINITE, BUF_WAYPOINT_INDEX = 5, 6, 7, 8
BUF_REC_DET_MASK, BUF_REC_DET_VALUES, BUF_REC_COLL_MASK, BUF_REC_LIDAR_HIT, BUF_REC_LIDAR_POINTS, BUF_REC_OMNI = 16, 17, 18, 19, 20, 21
PRECISIONS = {"fp32": PMS_FP32, "mixed": PMS_MIXED, "fp64": PMS_FP64, "fp32_plain": PMS_FP32_PLAIN}


class PmsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libpms_b200 error {code}: {msg}")
        self.code = code


class HsfmParamsC(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("tau", "A", "B", "k_1", "k_2", "k_o", "k_d", "k_lambda", "alpha",
                                           "pedestrian_mass", "pedestrian_radius", "robot_radius")]


class OrcaParamsC(C.Structure):
    _fields_ = [("time_step", C.c_double), ("neighbor_dist", C.c_double), ("max_neighbors", C.c_int),
                ("time_horizon", C.c_double), ("radius", C.c_double), ("goal_reaching_threshold", C.c_double)]


class LaunchInfoC(C.Structure):
    _fields_ = [("kernel", C.c_int), ("threads_per_block", C.c_int), ("blocks", C.c_int),
                ("worlds_per_block", C.c_int), ("j_split", C.c_int), ("smem_bytes", C.c_int),
                ("n_chunks", C.c_int), ("blocks_per_sm", C.c_int), ("kernel_launches", C.c_int64)]


_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)
_up = C.POINTER(C.c_uint32)
_lp = C.POINTER(C.c_int64)
_h = C.c_void_p

# name -> (restype, argtypes); every symbol include/pms.h declares
PROTOTYPES = {
    "pms_last_error": (C.c_char_p, []),
    "pms_version": (C.c_int, []),
    "pms_default_hsfm_params": (None, [C.POINTER(HsfmParamsC)]),
    "pms_default_orca_params": (None, [C.POINTER(OrcaParamsC)]),
    "pms_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(HsfmParamsC), C.POINTER(_h)]),
    "pms_destroy": (C.c_int, [_h]),
    "pms_n_worlds": (C.c_int, [_h]),
    "pms_n_peds": (C.c_int, [_h]),
    "pms_set_ped_state": (C.c_int, [_h, _dp, _dp]),
    "pms_get_ped_state": (C.c_int, [_h, _dp, _dp]),
    "pms_set_ped_state_f32": (C.c_int, [_h, _fp, _fp]),
    "pms_get_ped_state_f32": (C.c_int, [_h, _fp, _fp]),
    "pms_set_ped_state_device": (C.c_int, [_h, C.c_void_p, C.c_void_p, C.c_int]),
    "pms_get_ped_state_device": (C.c_int, [_h, C.c_void_p, C.c_void_p, C.c_int]),
    "pms_device_buffer": (C.c_int, [_h, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "pms_get_detector_f32": (C.c_int, [_h, _up, _fp]),
    "pms_set_ped_speeds": (C.c_int, [_h, _dp]),
    "pms_set_async_io": (C.c_int, [_h, C.c_int]),
    "pms_snapshot_save": (C.c_int, [_h]),
    "pms_snapshot_restore": (C.c_int, [_h]),
    "pms_set_waypoints_fixed": (C.c_int, [_h, _dp, C.c_int, _ip, C.c_double, C.c_int]),
    "pms_set_waypoint_queue": (C.c_int, [_h, _dp, _dp, C.c_int, C.c_double]),
    "pms_append_waypoints": (C.c_int, [_h, C.c_int, _ip, _dp]),
    "pms_pop_waypoint_events": (C.c_int, [_h, _ip, C.c_int, C.POINTER(C.c_int)]),
    "pms_get_waypoints": (C.c_int, [_h, _dp, _ip]),
    "pms_set_robot": (C.c_int, [_h, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int]),
    "pms_set_robot_controls": (C.c_int, [_h, _dp, C.c_int]),
    "pms_set_robot_controls_device": (C.c_int, [_h, C.c_void_p, C.c_int]),
    "pms_get_robot": (C.c_int, [_h, _dp, _dp, _ip]),
    "pms_set_map": (C.c_int, [_h, C.c_int, _dp, C.c_int, C.c_int]),
    "pms_set_accel_noise": (C.c_int, [_h, _dp]),
    "pms_set_accel_noise_std": (C.c_int, [_h, C.c_double, C.c_uint64]),
    "pms_set_lidar_noise": (C.c_int, [_h, C.c_int, C.c_double, _dp, _dp, C.c_uint64]),
    "pms_set_omni_noise": (C.c_int, [_h, C.c_int, C.c_double, C.c_double, C.c_double, C.c_uint64]),
    "pms_philox4x32_10": (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "pms_detector_config": (C.c_int, [_h, C.c_int, C.c_double, C.c_double, C.c_int]),
    "pms_set_detector_noise": (C.c_int, [_h, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint64]),
    "pms_hsfm_step": (C.c_int, [_h, C.c_double, C.c_int]),
    "pms_hsfm_step_async": (C.c_int, [_h, C.c_double, C.c_int]),
    "pms_sync": (C.c_int, [_h]),
    "pms_sense_tick": (C.c_int, [_h, _dp, _dp, C.c_double, _up, _up, _dp]),
    "pms_step_tick": (C.c_int, [_h, C.c_double, C.c_int, _dp, _dp, C.c_double, C.c_int, _dp, _dp, _ip, _up, _up, _dp, _ip, _ip]),
    "pms_stream": (C.c_void_p, [_h]),
    "pms_sense": (C.c_int, [_h]),
    "pms_get_detector": (C.c_int, [_h, _up, _dp]),
    "pms_get_collisions": (C.c_int, [_h, _up]),
    "pms_get_stats": (C.c_int, [_h, _lp, _lp, _ip]),
    "pms_reset_stats": (C.c_int, [_h]),
    "pms_lidar_scan": (C.c_int, [_h, C.c_int, C.c_double, C.c_double, _ip, _dp]),
    "pms_omni_scan": (C.c_int, [_h, _dp]),
    "pms_set_replay": (C.c_int, [_h, _dp, _dp, C.c_int, C.c_int]),
    "pms_replay_step": (C.c_int, [_h, C.c_double, C.c_int]),
    "pms_replay_frame": (C.c_int, [_h, C.c_int, C.POINTER(C.c_int)]),
    "pms_replay_tick": (C.c_int, [_h, C.c_int, _dp, C.c_double, _up, _up, _dp, _ip]),
    "pms_lidar_config": (C.c_int, [_h, C.c_int, C.c_double, C.c_double]),
    "pms_sensor_schedule": (C.c_int, [_h, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]),
    "pms_get_sensor_fires": (C.c_int, [_h, C.c_int, _ip, C.c_int, C.POINTER(C.c_int), _dp]),
    "pms_get_detector_records": (C.c_int, [_h, C.c_int, C.c_int, _up, _dp]),
    "pms_get_collision_records": (C.c_int, [_h, C.c_int, C.c_int, _up]),
    "pms_get_lidar_records": (C.c_int, [_h, C.c_int, C.c_int, _ip, _dp]),
    "pms_get_omni_records": (C.c_int, [_h, C.c_int, C.c_int, _dp]),
    "pms_orca_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(OrcaParamsC), C.POINTER(_h)]),
    "pms_orca_set_agents": (C.c_int, [_h, _dp, _dp, _dp, _dp]),
    "pms_orca_set_waypoints_fixed": (C.c_int, [_h, _dp, C.c_int, _ip, C.c_double, C.c_int]),
    "pms_orca_set_waypoints_current": (C.c_int, [_h, _dp]),
    "pms_orca_set_pref_refresh": (C.c_int, [_h, C.c_int]),
    "pms_orca_step": (C.c_int, [_h, C.c_int]),
    "pms_orca_step_async": (C.c_int, [_h, C.c_int]),
    "pms_orca_snapshot_save": (C.c_int, [_h]),
    "pms_orca_snapshot_restore": (C.c_int, [_h]),
    "pms_orca_get_agents": (C.c_int, [_h, _dp, _dp]),
    "pms_orca_tick": (C.c_int, [_h, C.c_int, _dp, _dp]),
    "pms_get_launch_info": (C.c_int, [_h, C.POINTER(LaunchInfoC)]),
    "pms_set_tuning": (C.c_int, [_h, C.c_int, C.c_int, C.c_int, C.c_int]),
    "pms_measure_fp32_tflops": (C.c_double, [C.c_int]),
    "pms_host_alloc": (C.c_void_p, [C.c_uint64]),
    "pms_host_free": (None, [C.c_void_p]),
}

_lib = None


def load():
    """Load libpms_b200.so.  Raises if it has not been built (python -m pyminisim_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -m pyminisim_b200.build` "
                              "(nvcc, sm_100a). pyminisim_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        raise PmsError(rc, load().pms_last_error().decode("utf-8", "replace"))


def dptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def fptr(a):
    return None if a is None else a.ctypes.data_as(_fp)


def iptr(a):
    return None if a is None else a.ctypes.data_as(_ip)


def uptr(a):
    return None if a is None else a.ctypes.data_as(_up)


def lptr(a):
    return None if a is None else a.ctypes.data_as(_lp)


def f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        a = np.ascontiguousarray(np.broadcast_to(a, shape))
    return a


def pinned_array(shape, dtype=np.float64):
    """numpy array backed by page-locked host memory from pms_host_alloc."""
    lib = load()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib.pms_host_alloc(n)
    if not p:
        raise MemoryError("pms_host_alloc failed")
    buf = (C.c_char * n).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    _PINNED.append((buf, p))  # pinned buffers live for the process lifetime
    return arr


_PINNED = []
