"""Drop-in mirror of ``pyminisim.sensors``: PedestrianDetector, LidarSensor, OmniObstacleDetector.

Same configs, readings, names ("pedestrian_detector", "lidar", "omni_obstacle_detector") and noise
models as the reference; the geometry runs on the GPU (pms_sense / pms_lidar_scan / pms_omni_scan).
Noise is drawn on the host from the global numpy RNG in the reference's order.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np

from . import _capi as capi
from .batched import BatchedSimulation
from .core import (AbstractSensor, AbstractSensorConfig, AbstractSensorReading, AbstractWorldMap, WorldState,
                   _GpuSenseContext)


# ---- pedestrian detector (sensors/_pedestrian_detector.py) ----------------------------------------------
@dataclass
class PedestrianDetectorNoise:
    distance_mu: float
    distance_sigma: float
    angle_mu: float
    angle_sigma: float
    misdetection_prob: float


class PedestrianDetectorConfig(AbstractSensorConfig):
    RETURN_POLAR = "polar"
    RETURN_ABSOLUTE = "absolute"
    RETURN_RELATIVE = "relative"
    RETURN_RELATIVE_ROTATED = "relative_rotated"
    _CODES = {"polar": capi.DET_POLAR, "relative_rotated": capi.DET_RELATIVE_ROTATED,
              "relative": capi.DET_RELATIVE, "absolute": capi.DET_ABSOLUTE}

    def __init__(self, max_dist: float, fov: float, return_type: str):
        assert return_type in self._CODES
        self._max_dist, self._fov, self._return_type = max_dist, fov, return_type

    max_dist = property(lambda self: self._max_dist)
    fov = property(lambda self: self._fov)
    return_type = property(lambda self: self._return_type)


class PedestrianDetectorReading(AbstractSensorReading):
    def __init__(self, poses: Dict[int, Tuple[float, float]]):
        self._pedestrians = poses

    @property
    def pedestrians(self) -> Dict[int, Tuple[float, float]]:
        return self._pedestrians


class PedestrianDetector(AbstractSensor):
    NAME = "pedestrian_detector"

    def __init__(self, config: PedestrianDetectorConfig = PedestrianDetectorConfig(max_dist=3., fov=np.deg2rad(30.),
                                                                                  return_type=PedestrianDetectorConfig.RETURN_POLAR),
                 noise: Optional[PedestrianDetectorNoise] = None):
        if noise is not None:
            assert noise.misdetection_prob <= 1.
        super().__init__(PedestrianDetector.NAME, period=0.)
        self._config, self._noise = config, noise
        self._ctx = _GpuSenseContext()

    @property
    def sensor_config(self) -> AbstractSensorConfig:
        return self._config

    def _fused_config(self):
        """(max_dist, fov, device return code) for a pedestrians model that evaluates the reading inside its step."""
        code = capi.DET_POLAR if self._noise is not None else self._config._CODES[self._config.return_type]
        return self._config.max_dist, self._config.fov, code

    def get_reading(self, world_state: WorldState, world_map: AbstractWorldMap) -> AbstractSensorReading:
        if world_state.pedestrians is None or world_state.robot is None:
            return PedestrianDetectorReading({})
        robot_pose = world_state.robot.pose
        # with noise the device returns polar readings; noise and the frame conversion follow on the host
        code = capi.DET_POLAR if self._noise is not None else self._config._CODES[self._config.return_type]
        # the fused step may have evaluated this very reading already (same robot pose, same configuration)
        fs = getattr(world_state.pedestrians, "_fused_sense", None)
        if fs is not None and fs[1] == (float(self._config.max_dist), float(self._config.fov), int(code)) \
                and np.array_equal(fs[0], robot_pose):
            hits = [(k, fs[3][k]) for k in fs[2]]
        else:
            ids, _, det, vals = self._ctx.run(world_state.pedestrians.poses, robot_pose, 0.35,
                                              (self._config.max_dist, self._config.fov, code))
            hits = [(k, v) for k, hit, v in zip(ids, det, vals) if hit]
        readings = {}
        for k, v in hits:
            if self._noise is None:
                readings[k] = (float(v[0]), float(v[1]))
                continue
            if not bool(np.random.binomial(1, 1. - self._noise.misdetection_prob)):     # :108
                continue
            distance = v[0] + np.random.normal(self._noise.distance_mu, self._noise.distance_sigma)
            angle = v[1] + np.random.normal(self._noise.angle_mu, self._noise.angle_sigma)
            readings[k] = self._convert(distance, angle, robot_pose)
        return PedestrianDetectorReading(readings)

    def _convert(self, distance, angle, robot_pose):                                    # :114-134
        rt = self._config.return_type
        if rt == PedestrianDetectorConfig.RETURN_POLAR:
            return distance, angle
        xr, yr = distance * np.cos(angle), distance * np.sin(angle)
        if rt == PedestrianDetectorConfig.RETURN_RELATIVE_ROTATED:
            return xr, yr
        th = robot_pose[2]
        x, y = xr * np.cos(th) - yr * np.sin(th), xr * np.sin(th) + yr * np.cos(th)
        if rt == PedestrianDetectorConfig.RETURN_RELATIVE:
            return x, y
        if rt == PedestrianDetectorConfig.RETURN_ABSOLUTE:
            return x + robot_pose[0], y + robot_pose[1]
        raise ValueError(f"Unknown return type {rt}")


# ---- shared one-world map context for lidar / omni --------------------------------------------------------
class _GpuMapContext:
    def __init__(self):
        self._sim = BatchedSimulation(1, 1, precision="fp64")
        self._sim.set_state(np.zeros((1, 1, 3)))
        self._map_id = None

    def place(self, world_map: AbstractWorldMap, robot_pose: np.ndarray) -> BatchedSimulation:
        kind = getattr(world_map, "map_kind", None)
        if kind is None:
            raise TypeError("the accelerated sensors support EmptyWorld, CirclesWorld and AABBWorld")
        key = (id(world_map), id(world_map.current_state))
        if key != self._map_id:
            self._sim.set_map(kind, world_map.device_data())
            self._map_id = key
        self._sim.set_robot(capi.ROBOT_EXTERNAL, np.asarray(robot_pose, dtype=np.float64)[None], velocity=np.zeros((1, 3)))
        return self._sim


# ---- lidar (sensors/_lidar.py) -----------------------------------------------------------------------------
class LidarSensorConfig(AbstractSensorConfig):
    def __init__(self, n_beams: int = 100, max_dist: float = 8., resolution: float = 0.01, frequency: float = 10.):
        assert n_beams > 0
        assert max_dist > resolution > 0.
        assert frequency > 0.
        self._n_beams, self._max_dist, self._resolution, self._frequency = n_beams, max_dist, resolution, frequency

    n_beams = property(lambda self: self._n_beams)
    max_dist = property(lambda self: self._max_dist)
    resolution = property(lambda self: self._resolution)
    frequency = property(lambda self: self._frequency)


class LidarSensorReading(AbstractSensorReading):
    def __init__(self, points: np.ndarray):
        self._points = points.copy()

    @property
    def points(self) -> np.ndarray:
        return self._points.copy()


class LidarSensorNoise:
    def __init__(self, point_mean: np.ndarray = np.array([0., 0.]), point_cov: np.ndarray = np.diag([0.001, 0.001]),
                 drop_prob: float = 0.):
        assert point_mean.shape == (2,)
        assert point_cov.shape == (2, 2)
        assert (point_cov >= 0.).all()
        assert 0. <= drop_prob <= 1.
        self._mean, self._cov, self._drop_prob = point_mean, point_cov, drop_prob

    mean = property(lambda self: self._mean.copy())
    cov = property(lambda self: self._cov.copy())
    drop_prob = property(lambda self: self._drop_prob)

    def noisify_reading(self, reading: LidarSensorReading) -> LidarSensorReading:
        pts = reading.points
        keep = np.random.binomial(1, 1 - self._drop_prob, pts.shape[0]).astype(bool)
        pts = pts[keep]
        if len(pts) == 0:
            return LidarSensorReading(np.array([]))
        return LidarSensorReading(pts + np.random.multivariate_normal(self._mean, self._cov, pts.shape[0]))


class LidarSensor(AbstractSensor):
    NAME = "lidar"

    def __init__(self, config: LidarSensorConfig = LidarSensorConfig(), noise: Optional[LidarSensorNoise] = None):
        period = 1. / config.frequency if np.isfinite(config.frequency) else 0.
        super().__init__(LidarSensor.NAME, period=period)
        self._config, self._noise = config, noise
        self._ctx = None

    @property
    def sensor_config(self) -> AbstractSensorConfig:
        return self._config

    def get_reading(self, world_state: WorldState, world_map: AbstractWorldMap) -> AbstractSensorReading:
        if world_state.robot is None:
            return LidarSensorReading(np.array([]))
        if self._ctx is None:
            self._ctx = _GpuMapContext()
        sim = self._ctx.place(world_map, world_state.robot.pose)
        hit, pts = sim.lidar_scan(self._config.n_beams, self._config.max_dist, self._config.resolution)
        hits = pts[0][hit[0] >= 0]                                  # beam order, misses dropped (:117-120)
        # no beam hit anything: the reference's list of points is empty and np.array([]) has shape (0,), not (0, 2) (:121)
        reading = LidarSensorReading(hits if hits.shape[0] else np.array([]))
        if self._noise is not None:
            reading = self._noise.noisify_reading(reading)
        return reading


# ---- omni obstacle detector (sensors/_omni_obstacle_detector.py) ----------------------------------------------
@dataclass
class OmniObstacleDetectorNoise:
    distance_mu: float
    distance_sigma: float
    misdetection_prob: float


class OmniObstacleDetectorConfig(AbstractSensorConfig):
    def __init__(self, max_dist: float = np.inf):
        self._max_dist = max_dist

    max_dist = property(lambda self: self._max_dist)


class OmniObstacleDetectorReading(AbstractSensorReading):
    def __init__(self, min_obstacle_dist: float):
        self._min_obstacle_dist = min_obstacle_dist

    @property
    def min_obstacle_dist(self) -> float:
        return self._min_obstacle_dist

    @property
    def is_finite(self) -> bool:
        return np.isfinite(self._min_obstacle_dist)


class OmniObstacleDetector(AbstractSensor):
    NAME = "omni_obstacle_detector"

    def __init__(self, config: OmniObstacleDetectorConfig = OmniObstacleDetectorConfig(),
                 noise: Optional[OmniObstacleDetectorNoise] = None):
        if noise is not None:
            assert noise.misdetection_prob <= 1.
        super().__init__(OmniObstacleDetector.NAME, period=0.)
        self._config, self._noise = config, noise
        self._ctx = None

    @property
    def sensor_config(self) -> AbstractSensorConfig:
        return self._config

    def get_reading(self, world_state: WorldState, world_map: AbstractWorldMap) -> AbstractSensorReading:
        if world_state.robot is None:
            return OmniObstacleDetectorReading(np.inf)
        if self._ctx is None:
            self._ctx = _GpuMapContext()
        dist = float(self._ctx.place(world_map, world_state.robot.pose).omni_scan()[0])
        if self._noise is not None:
            if not bool(np.random.binomial(1, 1. - self._noise.misdetection_prob)):
                dist = np.inf
            else:
                dist = dist + np.random.normal(self._noise.distance_mu, self._noise.distance_sigma)
        return OmniObstacleDetectorReading(dist)


# ---------------------------------------------------------------------------------------------------------
# SemanticDetector (sensors/_semantic_object_detector.py): named AABB objects near the robot.  Host arithmetic over a
# handful of boxes (AABBWorld.closest_semantic_objects) -- kept for drop-in completeness, not on the accelerated path.
# ---------------------------------------------------------------------------------------------------------
class SemanticDetectorConfig(AbstractSensorConfig):
    def __init__(self, max_dist: float = 8., frequency: float = 10.):
        assert max_dist > 0.
        assert frequency > 0.
        super().__init__()
        self._max_dist, self._frequency = max_dist, frequency

    max_dist = property(lambda self: self._max_dist)
    frequency = property(lambda self: self._frequency)


@dataclass
class SemanticDetection:
    class_name: str
    object_name: Optional[str]
    position: np.ndarray


class SemanticDetectorReading(AbstractSensorReading):
    def __init__(self, detections: Dict[int, SemanticDetection]):
        super().__init__()
        self._detections = detections

    @property
    def detections(self) -> Dict[int, SemanticDetection]:
        return self._detections


class SemanticDetectorNoise:
    def __init__(self, position_mean: np.ndarray = np.array([0., 0.]), position_cov: np.ndarray = np.diag([0.001, 0.001]),
                 drop_prob: float = 0.):
        assert position_mean.shape == (2,)
        assert position_cov.shape == (2, 2)
        assert (position_cov >= 0.).all()
        assert 0. <= drop_prob <= 1.
        self._position_mean, self._position_cov, self._drop_prob = position_mean, position_cov, drop_prob

    mean = property(lambda self: self._position_mean.copy())
    cov = property(lambda self: self._position_cov.copy())
    drop_prob = property(lambda self: self._drop_prob)

    def noisify_reading(self, reading: SemanticDetectorReading) -> SemanticDetectorReading:
        # the global numpy generator, one binomial then one multivariate normal draw per reading (:73-82); a detection
        # survives where the binomial draw (probability 1 - drop_prob) is ZERO -- the reference's rule, reproduced
        found = reading.detections
        drawn = np.random.binomial(1, 1 - self._drop_prob, len(found)).astype(bool)
        shift = np.random.multivariate_normal(self._position_mean, self._position_cov, len(found))
        return SemanticDetectorReading({idx: SemanticDetection(d.class_name, d.object_name, d.position + shift[k])
                                        for k, (idx, d) in enumerate(found.items()) if not drawn[k]})


class SemanticDetector(AbstractSensor):
    NAME = "semantic_detector"

    def __init__(self, config: SemanticDetectorConfig = SemanticDetectorConfig(), noise: Optional[SemanticDetectorNoise] = None):
        super().__init__(SemanticDetector.NAME, period=1. / config.frequency if np.isfinite(config.frequency) else 0.)
        self._config, self._noise = config, noise

    @property
    def sensor_config(self) -> AbstractSensorConfig:
        return self._config

    def get_reading(self, world_state: WorldState, world_map: AbstractWorldMap) -> AbstractSensorReading:
        from .world_map import AABBWorld
        assert isinstance(world_map, AABBWorld), \
            f"Semantic detector can work only with AABBWorld worlds, got instance of {type(world_map)}"
        near = world_map.closest_semantic_objects(world_state.robot.pose[:2], self._config.max_dist)
        reading = SemanticDetectorReading({idx: SemanticDetection(*v) for idx, v in near.items()})
        return self._noise.noisify_reading(reading) if self._noise is not None else reading
