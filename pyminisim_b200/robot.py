"""Drop-in mirror of ``pyminisim.robot``: unicycle and bicycle kinematics (robot/_unicycle.py, _bicycle.py).

In the single-world drop-in these run on the host exactly like the reference (an O(1) explicit-Euler
update whose result is an *input* of the hot path); in the batched API the same models run inside the
fused kernel's robot warp (csrc/hsfm_small.cuh: robot_step)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .core import (DEFAULT_ROBOT_RADIUS, AbstractRobotMotionModel, AbstractRobotMotionModelState, AbstractWorldMap)


def _wrap(theta: float) -> float:
    return (theta + np.pi) % (2 * np.pi) - np.pi


class UnicycleRobotModelState(AbstractRobotMotionModelState):
    def __init__(self, state: np.ndarray, control: np.ndarray):
        assert state.shape == (3,)
        assert control.shape == (2,)
        self._state, self._control = state.copy(), control.copy()

    @property
    def pose(self) -> np.ndarray:
        return self._state[:3].copy()

    @property
    def velocity(self) -> np.ndarray:
        v, w, theta = self._control[0], self._control[1], self._state[2]
        return np.array([v * np.cos(theta), v * np.sin(theta), w])

    @property
    def state(self) -> np.ndarray:
        return self._state.copy()

    @property
    def control(self) -> np.ndarray:
        return self._control.copy()


class UnicycleRobotModel(AbstractRobotMotionModel):
    def __init__(self, initial_pose: np.ndarray, initial_control: np.ndarray = np.array([0., 0.]),
                 robot_radius: float = DEFAULT_ROBOT_RADIUS):
        super().__init__(UnicycleRobotModelState(initial_pose, initial_control), robot_radius)

    def step(self, dt: float, world_map: AbstractWorldMap, control: Optional[np.ndarray] = None) -> bool:
        if control is None:
            control = self._state.control
        pose = self._state.pose
        x = pose[0] + control[0] * np.cos(pose[2]) * dt
        y = pose[1] + control[0] * np.sin(pose[2]) * dt
        theta = _wrap(pose[2] + control[1] * dt)
        if world_map.is_collision(np.array([x, y]), self._radius):
            return False           # move vetoed: pose and control stay
        self._state = UnicycleRobotModelState(np.array([x, y, theta]), np.array([control[0], control[1]]))
        return True


class BicycleRobotModelState(AbstractRobotMotionModelState):
    def __init__(self, center_pose: np.ndarray, rear_pose: np.ndarray, control: np.ndarray, wheel_base: float):
        assert center_pose.shape == (3,)
        assert rear_pose.shape == (3,)
        assert control.shape == (2,)
        self._center, self._rear, self._control, self._wheel_base = center_pose.copy(), rear_pose.copy(), control.copy(), wheel_base

    @property
    def pose(self) -> np.ndarray:
        return self._center.copy()

    @property
    def velocity(self) -> np.ndarray:
        return np.zeros((3,))      # the reference reports zero (robot/_bicycle.py:30-32)

    @property
    def state(self) -> np.ndarray:
        return self._center.copy()

    @property
    def control(self) -> np.ndarray:
        return self._control.copy()

    @property
    def wheel_base(self) -> float:
        return self._wheel_base

    @property
    def rear_pose(self) -> np.ndarray:
        return self._rear.copy()


class BicycleRobotModel(AbstractRobotMotionModel):
    def __init__(self, wheel_base: float, initial_center_pose: np.ndarray,
                 initial_control: np.ndarray = np.array([0., 0.]), robot_radius: float = DEFAULT_ROBOT_RADIUS):
        c = initial_center_pose
        rear = np.array([c[0] - wheel_base / 2. * np.cos(c[2]), c[1] - wheel_base / 2. * np.sin(c[2]), c[2]])
        super().__init__(BicycleRobotModelState(c, rear, initial_control, wheel_base), robot_radius)
        self._l = wheel_base

    @property
    def wheel_base(self) -> float:
        return self._l

    def step(self, dt: float, world_map: AbstractWorldMap, control: Optional[np.ndarray] = None) -> bool:
        if control is None:
            control = self._state.control
        v, delta = control[0], control[1]
        rear = self._state.rear_pose
        xr = rear[0] + v * np.cos(rear[2]) * dt
        yr = rear[1] + v * np.sin(rear[2]) * dt
        theta = _wrap(rear[2] + (v / self._l) * np.tan(delta) * dt)
        xc, yc = xr + self._l / 2. * np.cos(theta), yr + self._l / 2. * np.sin(theta)
        if world_map.is_collision(np.array([xc, yc]), self._radius):
            return False
        self._state = BicycleRobotModelState(np.array([xc, yc, theta]), np.array([xr, yr, theta]), np.array([v, delta]), self._l)
        return True
