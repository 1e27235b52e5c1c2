"""Drop-in mirror of ``pyminisim.core`` for the accelerated path (reference: pyminisim/core/*.py).

Interfaces (names, argument meaning, error behaviour) follow the reference so that user code written
against ``pyminisim.core`` keeps working; the collision list and every sensor reading are computed by
libpms_b200 on the GPU.  Only what the hot path and its callers need is here -- no renderer.
"""
from __future__ import annotations

import random
import time
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

DEFAULT_ROBOT_RADIUS: float = 0.35   # core/_constants.py:1
PEDESTRIAN_RADIUS: float = 0.3       # core/_constants.py:2


# ---------------------------------------------------------------------------------------------------------------
# The abstract interfaces and value types of pyminisim.core.  When the reference package itself is importable its OWN
# classes are used, so everything in this package subclasses the reference's ABCs (isinstance checks, type hints and the
# reference's own Simulation / renderer accept the accelerated models and sensors).  Otherwise -- the GPU box has no
# reference -- the definitions below stand in for them: same names, members and assertion behaviour.
# ---------------------------------------------------------------------------------------------------------------
def _load_reference_core():
    import importlib
    import sys
    try:
        pkg = importlib.import_module("pyminisim")
        if getattr(pkg, "__pms_b200_alias__", False):       # our own alias (pyminisim_b200.alias), not the reference
            return None
        core = importlib.import_module("pyminisim.core")
        sensor = importlib.import_module("pyminisim.core._sensor")
        return core, sensor
    except Exception:
        for name in [m for m in sys.modules if m == "pyminisim" or m.startswith("pyminisim.")]:
            if not getattr(sys.modules.get("pyminisim"), "__pms_b200_alias__", False):
                sys.modules.pop(name, None)                   # a half-imported reference must not linger
        return None


_REF = _load_reference_core()
USING_REFERENCE_ABCS = _REF is not None

if _REF is not None:
    _c, _s = _REF
    AbstractWorldMapState, AbstractWorldMap, AbstractStaticWorldMap = _c.AbstractWorldMapState, _c.AbstractWorldMap, _c.AbstractStaticWorldMap
    AbstractRobotMotionModelState, AbstractRobotMotionModel = _c.AbstractRobotMotionModelState, _c.AbstractRobotMotionModel
    AbstractWaypointTrackerState, AbstractWaypointTracker = _c.AbstractWaypointTrackerState, _c.AbstractWaypointTracker
    AbstractPedestriansModelState, AbstractPedestriansModel = _c.AbstractPedestriansModelState, _c.AbstractPedestriansModel
    AbstractSensorConfig, AbstractSensorReading, AbstractSensor = _c.AbstractSensorConfig, _c.AbstractSensorReading, _c.AbstractSensor
    SensorState, WorldState, SimulationState = _s.SensorState, _c.WorldState, _c.SimulationState
else:
    # ---- world map (core/_world_map.py:7-71) ------------------------------------------------------------
    class AbstractWorldMapState(ABC):
        pass


    class AbstractWorldMap(ABC):
        def __init__(self, initial_state: AbstractWorldMapState):
            self._state = initial_state

        @property
        def current_state(self) -> AbstractWorldMapState:
            return self._state

        def reset_to_state(self, state: AbstractWorldMapState):
            self._state = state

        @abstractmethod
        def step(self, dt: float) -> None: ...

        @abstractmethod
        def closest_distance_to_obstacle(self, point: np.ndarray) -> Union[float, np.ndarray]: ...

        @abstractmethod
        def is_occupied(self, point: np.ndarray) -> Union[bool, np.ndarray]: ...

        @abstractmethod
        def is_collision(self, agent_points: np.ndarray, agent_radii, eps_margin: float = 0.05): ...


    class AbstractStaticWorldMap(AbstractWorldMap, ABC):
        def step(self, dt: float) -> None:
            return


    # ---- robot (core/_motion_state.py:6-42, core/_motion.py:10-32) ---------------------------------------
    class AbstractRobotMotionModelState(ABC):
        @property
        @abstractmethod
        def pose(self) -> np.ndarray: ...

        @property
        @abstractmethod
        def velocity(self) -> np.ndarray: ...

        @property
        @abstractmethod
        def state(self) -> np.ndarray: ...

        @property
        @abstractmethod
        def control(self) -> np.ndarray: ...


    class AbstractRobotMotionModel(ABC):
        def __init__(self, initial_state: AbstractRobotMotionModelState, radius: float = DEFAULT_ROBOT_RADIUS):
            assert radius > 0.
            self._state = initial_state
            self._radius = radius

        @property
        def state(self) -> AbstractRobotMotionModelState:
            return self._state

        @property
        def radius(self) -> float:
            return self._radius

        def reset_to_state(self, state: AbstractRobotMotionModelState):
            self._state = state

        @abstractmethod
        def step(self, dt: float, world_map: AbstractWorldMap, control: Optional[np.ndarray] = None) -> bool: ...


    # ---- waypoints (core/_waypoints.py:7-40) ---------------------------------------------------------------
    class AbstractWaypointTrackerState(ABC):
        def __init__(self, current_waypoints: Dict[int, np.ndarray]):
            self._current_waypoints = {}
            for ped_id, wp in current_waypoints.items():
                assert wp.shape == (2,), f"Waypoints must be array of shape (2,), the {wp.shape} array was passed for pedestrian {ped_id}"
                self._current_waypoints[ped_id] = wp.copy()

        @property
        def current_waypoints(self) -> Dict[int, np.ndarray]:
            return self._current_waypoints


    class AbstractWaypointTracker(ABC):
        @property
        def state(self) -> AbstractWaypointTrackerState:
            raise NotImplementedError()

        @abstractmethod
        def resample_all(self, agents_poses: Dict[int, np.ndarray]) -> Dict[int, np.ndarray]: ...

        @abstractmethod
        def update_waypoints(self, agents_poses: Dict[int, np.ndarray]) -> Dict[int, Tuple[np.ndarray, bool]]: ...

        @abstractmethod
        def reset_to_state(self, state: AbstractWaypointTrackerState): ...

        @abstractmethod
        def sample_independent_points(self, n_points: int, min_cross_distance: Optional[float] = None): ...


    # ---- pedestrians (core/_pedestrians_model.py:9-47) -------------------------------------------------------
    class AbstractPedestriansModelState(ABC):
        def __init__(self, pedestrians: Dict[int, Tuple[np.ndarray, np.ndarray]],
                     waypoints_state: Optional[AbstractWaypointTrackerState]):
            for ped_id, (pose, velocity) in pedestrians.items():
                assert pose.shape == (3,), f"Each pedestrian pose must be a 3-dim array, but {pose.shape} array was passed for pedestrian {ped_id}"
                assert velocity.shape == (3,), f"Each pedestrian velocity must be a 3-dim array, but {velocity.shape} array was passed for pedestrian {ped_id}"
            self._pedestrians = pedestrians
            self._waypoints = waypoints_state

        @property
        def poses(self) -> Dict[int, np.ndarray]:
            return {k: v[0].copy() for k, v in self._pedestrians.items()}      # fresh copies, like the reference

        @property
        def velocities(self) -> Dict[int, np.ndarray]:
            return {k: v[1].copy() for k, v in self._pedestrians.items()}

        @property
        def waypoints(self) -> Optional[AbstractWaypointTrackerState]:
            return self._waypoints


    class AbstractPedestriansModel(ABC):
        @property
        @abstractmethod
        def state(self) -> AbstractPedestriansModelState: ...

        @abstractmethod
        def step(self, dt: float, robot_pose: Optional[np.ndarray], robot_velocity: Optional[np.ndarray]): ...

        @abstractmethod
        def reset_to_state(self, state: AbstractPedestriansModelState): ...


    # ---- sensors (core/_sensor.py:8-44) ----------------------------------------------------------------------
    class AbstractSensorConfig(ABC):
        pass


    class AbstractSensorReading(ABC):
        pass


    class AbstractSensor(ABC):
        def __init__(self, name: str, period: float):
            assert period >= 0.
            self._name = name
            self._period = period

        @property
        def sensor_name(self) -> str:
            return self._name

        @property
        def period(self) -> float:
            return self._period

        @property
        @abstractmethod
        def sensor_config(self) -> AbstractSensorConfig: ...

        @abstractmethod
        def get_reading(self, world_state: "WorldState", world_map: AbstractWorldMap) -> AbstractSensorReading: ...


    @dataclass
    class SensorState:
        reading: AbstractSensorReading
        hold_time: float


    # ---- state containers (core/_world_state.py:11-17, core/_simulation_state.py:10-13) -----------------------
    @dataclass
    class WorldState:
        world_map: AbstractWorldMapState
        robot: Optional[AbstractRobotMotionModelState]
        pedestrians: Optional[AbstractPedestriansModelState]
        robot_to_pedestrians_collisions: Optional[List[int]]
        robot_to_world_collision: bool


    @dataclass
    class SimulationState:
        world: WorldState
        sensors: Dict[str, SensorState]



class _GpuSenseContext:
    """One-world handle used for the robot-pedestrian collision list (core/_simulation.py:141-143) and by
    the pedestrian detector: uploads the poses, places the robot, runs the fused sense kernel."""

    def __init__(self):
        self._sim = None
        self._n = -1

    def run(self, ped_poses, robot_pose: np.ndarray, robot_radius: float,
            detector: Optional[Tuple[float, float, int]] = None):
        """ped_poses: the reference's ``{id: pose}`` dict, or an (N, 3) array of pedestrians 0..N-1 (array-backed states)."""
        from . import _capi as capi
        from .batched import BatchedSimulation
        as_array = isinstance(ped_poses, np.ndarray)
        ids = list(range(ped_poses.shape[0])) if as_array else list(ped_poses.keys())
        n = len(ids)
        if n == 0:
            return ids, np.zeros((0,), dtype=bool), np.zeros((0,), dtype=bool), np.zeros((0, 2))
        if self._sim is None or self._n != n:
            if self._sim is not None:
                self._sim.close()
            self._sim = BatchedSimulation(1, n, precision="fp64")
            self._n = n
        poses = ped_poses[None] if as_array else np.stack([ped_poses[k] for k in ids])[None]
        if detector is None:
            self._sim.set_detector(enabled=False)
        else:
            self._sim.set_detector(detector[0], detector[1], detector[2], enabled=True)
        # one call, one synchronisation: poses + robot in, collision list + reading out
        coll, mask, vals = self._sim.sense_tick(poses, np.asarray(robot_pose, dtype=np.float64)[None], robot_radius)
        k = np.arange(n)
        bits = lambda m: ((m[k >> 5] >> (k & 31).astype(np.uint32)) & 1).astype(bool)  # noqa: E731
        return ids, bits(coll[0]), bits(mask[0]), vals[0]


# ---- the orchestrator (core/_simulation.py:16-185) -----------------------------------------------------------
class Simulation:
    """Same step order as the reference: world map -> robot -> pedestrians (which see the robot's NEW pose
    and velocity) -> world state with the collision list -> sensors on their hold-time schedule."""

    def __init__(self, world_map: AbstractWorldMap, robot_model: Optional[AbstractRobotMotionModel],
                 pedestrians_model: Optional[AbstractPedestriansModel], sensors: List[AbstractSensor],
                 sim_dt: float = 0.01, rt_factor: Optional[float] = None):
        self._world_map = world_map
        self._robot_model = robot_model
        self._pedestrians_model = pedestrians_model
        self._sensors = sensors
        self._sim_dt = sim_dt
        self._rt_factor = rt_factor
        self._backup_robot_model = None
        self._backup_pedestrians_model = None
        self._robot_to_world_collision = False
        self._sense = _GpuSenseContext()
        self._detector_cfg = None           # (max_dist, fov, device return code) of the first pedestrian detector
        for sensor in sensors:
            if getattr(sensor, "sensor_name", None) == "pedestrian_detector" and hasattr(sensor, "_fused_config"):
                self._detector_cfg = sensor._fused_config()
                break
        if robot_model is not None and hasattr(pedestrians_model, "_collision_radius"):
            pedestrians_model._collision_radius = robot_model.radius      # the fused step evaluates the collision list with it
            if self._detector_cfg is not None and hasattr(pedestrians_model, "configure_fused_detector"):
                pedestrians_model.configure_fused_detector(*self._detector_cfg)   # ... and the detector's reading
        self._current_state = self._make_state({})       # the constructor runs the sensor pass once (:37)

    @property
    def sim_dt(self) -> float:
        return self._sim_dt

    @property
    def world_map(self) -> AbstractWorldMap:
        return self._world_map

    @property
    def current_state(self) -> SimulationState:
        return self._current_state

    @property
    def sensors(self) -> List[AbstractSensor]:
        return self._sensors

    @staticmethod
    def seed(seed: int) -> None:
        random.seed(seed)
        np.random.seed(seed)

    def step(self, control: Optional[np.ndarray] = None) -> SimulationState:
        t0 = time.time()
        self._advance(control)
        elapsed = time.time() - t0
        if self._rt_factor is not None and self._sim_dt / self._rt_factor - elapsed > 0.:
            time.sleep(self._sim_dt / self._rt_factor)    # the reference sleeps the full period (:66-69)
        self._current_state = self._make_state(self._current_state.sensors)
        return self._current_state

    def reset_to_state(self, state: WorldState):
        self._world_map.reset_to_state(state.world_map)
        if self._robot_model is not None:
            self._robot_model.reset_to_state(state.robot)
        if self._pedestrians_model is not None:
            self._pedestrians_model.reset_to_state(state.pedestrians)
        self._robot_to_world_collision = state.robot_to_world_collision
        self._current_state = self._make_state({})

    def _toggle(self, enabled: bool, live: str, backup: str, what: str):
        if enabled:
            if getattr(self, live) is not None:
                return
            if getattr(self, backup) is None:
                raise RuntimeError(f"{what} model was not initialized")
            setattr(self, live, getattr(self, backup))
            setattr(self, backup, None)
        else:
            if getattr(self, backup) is not None:
                return
            if getattr(self, live) is None:
                raise RuntimeError(f"{what} model was not initialized")
            setattr(self, backup, getattr(self, live))
            setattr(self, live, None)
        self._current_state = self._make_state(self._current_state.sensors)

    def set_robot_enabled(self, enabled: bool):
        self._toggle(enabled, "_robot_model", "_backup_robot_model", "Robot")

    def set_pedestrians_enabled(self, enabled: bool):
        self._toggle(enabled, "_pedestrians_model", "_backup_pedestrians_model", "Pedestrians")

    # -- internals --
    def _advance(self, control: Optional[np.ndarray]):
        self._world_map.step(self._sim_dt)
        robot_pose = robot_velocity = None
        if self._robot_model is not None:
            self._robot_to_world_collision = not self._robot_model.step(self._sim_dt, self._world_map, control)
            robot_pose = self._robot_model.state.pose
            robot_velocity = self._robot_model.state.velocity
        if self._pedestrians_model is not None:
            self._pedestrians_model.step(self._sim_dt, robot_pose, robot_velocity)

    def _world_state(self) -> WorldState:
        robot_state = self._robot_model.state if self._robot_model is not None else None
        ped_state = self._pedestrians_model.state if self._pedestrians_model is not None else None
        collisions = None
        if robot_state is not None and ped_state is not None:
            # the fused step already evaluated the list for this robot pose (HSFM policy); else a sense pass of its own
            fused = getattr(self._pedestrians_model, "fused_collisions", None)
            collisions = fused(robot_state.pose, self._robot_model.radius) if fused is not None else None
            if collisions is None:
                # one pass serves the collision list and the (first) pedestrian detector: its reading travels with the
                # state object like the fused step's does
                cfg = self._detector_cfg
                arrays = getattr(ped_state, "_arrays", None)        # array-backed state: no dict of copies for the upload
                ids, coll, det, vals = self._sense.run(arrays[0] if arrays is not None else ped_state.poses, robot_state.pose,
                                                       self._robot_model.radius, cfg)
                collisions = [k for k, c in zip(ids, coll) if c]
                if cfg is not None and ids == list(range(len(ids))):
                    try:
                        ped_state._fused_sense = (np.array(robot_state.pose, dtype=np.float64),
                                                  (float(cfg[0]), float(cfg[1]), int(cfg[2])),
                                                  [k for k, d in zip(ids, det) if d], vals)
                    except AttributeError:        # a foreign state class with __slots__: the detector runs its own pass
                        pass
        return WorldState(world_map=self._world_map.current_state, robot=robot_state, pedestrians=ped_state,
                          robot_to_pedestrians_collisions=collisions,
                          robot_to_world_collision=self._robot_to_world_collision)

    def _make_state(self, previous: Dict[str, SensorState]) -> SimulationState:
        world = self._world_state()
        return SimulationState(world=world, sensors=self._sensor_pass(world, previous))

    def _sensor_pass(self, world: WorldState, previous: Dict[str, SensorState]) -> Dict[str, SensorState]:
        """Hold-time scheduler (core/_simulation.py:162-185).  hold_time accumulates sim_dt in fp64, so a
        0.1 s period fires every 11 steps of 0.01 s (ten additions give 0.0999...), like the reference."""
        if self._robot_model is None:
            return {}
        out = {}
        for sensor in self._sensors:
            prev = previous.get(sensor.sensor_name) if len(previous) else None
            if prev is None:
                out[sensor.sensor_name] = SensorState(sensor.get_reading(world, self._world_map), 0.)
                continue
            hold = prev.hold_time + self._sim_dt
            if hold >= sensor.period:
                out[sensor.sensor_name] = SensorState(sensor.get_reading(world, self._world_map), 0.)
            else:
                out[sensor.sensor_name] = SensorState(prev.reading, hold)
        return out
