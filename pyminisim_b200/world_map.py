"""Drop-in mirror of ``pyminisim.world_map``: EmptyWorld, CirclesWorld, AABBWorld.

The maps are small host-side containers (the data format the lidar / omni / robot-veto kernels consume);
their geometric predicates are evaluated on the GPU by libpms_b200 (pms_set_map + pms_lidar_scan /
pms_omni_scan, and the veto inside the fused step when the robot model runs on the device).  The numpy
methods below reproduce the reference's results for callers that query a map directly.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

from . import _capi as capi
from .core import AbstractStaticWorldMap, AbstractWorldMapState


def _as_points(point: np.ndarray) -> Tuple[np.ndarray, bool]:
    assert point.shape == (2,) or (len(point.shape) == 2 and point.shape[1] == 2), \
        f"point must have shape (2,) or (N, 2), got {point.shape}"
    return (point[np.newaxis, :], True) if point.shape == (2,) else (point, False)


class _MapBase(AbstractStaticWorldMap):
    """Shared is_collision (core/_world_map.py:35-62) and the device-format accessors."""
    map_kind = capi.MAP_EMPTY

    def device_data(self) -> Optional[np.ndarray]:
        return None

    def is_collision(self, agent_points: np.ndarray, agent_radii, eps_margin: float = 0.05):
        assert eps_margin >= 0., f"eps_margin must be >= 0., got {eps_margin}"
        pts, single = _as_points(agent_points)
        radii = np.broadcast_to(np.asarray(agent_radii, dtype=np.float64), (pts.shape[0],))
        hit = np.logical_or(self.is_occupied(pts), self.closest_distance_to_obstacle(pts) - radii + eps_margin <= 0.)
        return hit[0] if single else hit


class EmptyWorldMapState(AbstractWorldMapState):
    pass


class EmptyWorld(_MapBase):
    def __init__(self):
        super().__init__(EmptyWorldMapState())

    def closest_distance_to_obstacle(self, point: np.ndarray):
        pts, single = _as_points(point)
        return np.inf if single else np.repeat(np.inf, pts.shape[0])

    def is_occupied(self, point: np.ndarray):
        return np.zeros(point.shape[0], dtype=bool) if len(point.shape) == 2 else False

    def is_collision(self, agent_points: np.ndarray, agent_radii, eps_margin: float = 0.05):
        if len(agent_points.shape) == 1:
            return False
        if len(agent_points.shape) == 2:
            return np.zeros(agent_points.shape[0], dtype=bool)
        raise ValueError(f"Unsupported agent_point shape: {agent_points.shape}")


class _CirclesWorldState(AbstractWorldMapState):
    def __init__(self, circles: np.ndarray):
        assert len(circles.shape) == 2 and circles.shape[1] == 3
        self._circles = circles

    @property
    def circles(self) -> np.ndarray:
        return self._circles


class CirclesWorld(_MapBase):
    """circles: rows of [x, y, radius] in metres (world_map/_circles_world.py:23-27)."""
    map_kind = capi.MAP_CIRCLES

    def __init__(self, circles: np.ndarray):
        super().__init__(_CirclesWorldState(np.array(circles, dtype=np.float64)))

    def device_data(self):
        return self._state.circles

    def _signed(self, pts: np.ndarray) -> np.ndarray:
        c = self._state.circles
        d = pts[:, None, :] - c[None, :, :2]
        return np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) - c[None, :, 2]

    def closest_distance_to_obstacle(self, point: np.ndarray):
        pts, single = _as_points(point)
        out = np.min(self._signed(pts), axis=1)
        return out[0] if single else out

    def is_occupied(self, point: np.ndarray):
        pts, single = _as_points(point)
        s = self._signed(pts)
        if single:   # the reference's single-point branch is inverted (True when free): _circles_world.py:44-45
            return bool((s[0] > 0.).all())
        return (s <= 0.).any(axis=1)

    @property
    def circles(self) -> np.ndarray:
        return self._state.circles.copy()


@dataclass
class AABBObject:
    box: Tuple[float, float, float, float]     # (tl_x, tl_y, width, height): x decreases by h, y increases by w
    class_name: Optional[str] = None
    object_name: Optional[str] = None
    color: Optional[Tuple[int, int, int]] = None


class _AABBWorldState(AbstractWorldMapState):
    def __init__(self, objects: List[AABBObject], default_color):
        boxes, segments, colors, semantic = [], [], [], {}
        for idx, obj in enumerate(objects):
            assert len(obj.box) == 4, f"Object box must be a tuple of 4 floats, got {obj.box} for object {idx}"
            if obj.object_name is not None and obj.class_name is None:
                raise ValueError(f"class_name must be specified if object_name is not None, violated for object at {idx}")
            tl_x, tl_y, w, h = obj.box
            boxes.append(obj.box)
            segments += [[tl_x, tl_y, tl_x, tl_y + w], [tl_x - h, tl_y, tl_x - h, tl_y + w],
                         [tl_x, tl_y, tl_x - h, tl_y], [tl_x, tl_y + w, tl_x - h, tl_y + w]]
            colors.append(obj.color if obj.color is not None else default_color)
            if obj.class_name is not None:
                semantic[idx] = (obj.class_name, obj.object_name)
        self._boxes = np.array(boxes, dtype=np.float64).reshape(-1, 4)
        self._segments = np.array(segments, dtype=np.float64).reshape(-1, 4)
        self._colors, self._semantic_objects = colors, semantic

    boxes = property(lambda self: self._boxes)
    segments = property(lambda self: self._segments)
    colors = property(lambda self: self._colors)
    semantic_objects = property(lambda self: self._semantic_objects)


class AABBWorld(_MapBase):
    map_kind = capi.MAP_AABB

    def __init__(self, objects: List[AABBObject], default_color=(0, 255, 0)):
        super().__init__(_AABBWorldState(objects, default_color))

    def device_data(self):
        return self._state.boxes

    def closest_distance_to_obstacle(self, point: np.ndarray):
        pts, single = _as_points(point)
        seg = self._state.segments
        p1, e = seg[:, :2], seg[:, 2:] - seg[:, :2]
        o = pts[:, None, :] - p1[None]
        t = np.clip(np.einsum("ijk,jk->ij", o, e) / np.einsum("ij,ij->i", e, e), 0, 1)
        q = p1[None] + t[..., None] * e[None]
        out = np.min(np.linalg.norm(pts[:, None, :] - q, axis=-1), axis=1)
        return out[0] if single else out

    def is_occupied(self, point: np.ndarray):
        pts, single = _as_points(point)
        res = np.zeros(pts.shape[0], dtype=bool)
        for tl_x, tl_y, w, h in self._state.boxes:
            res |= (pts[:, 0] <= tl_x) & (pts[:, 0] >= tl_x - h) & (pts[:, 1] >= tl_y) & (pts[:, 1] <= tl_y + w)
        return res[0] if single else res

    @property
    def boxes(self) -> np.ndarray:
        return self._state.boxes.copy()

    @property
    def colors(self):
        return self._state.colors

    def closest_semantic_objects(self, point: np.ndarray, max_distance: float) -> Dict[int, Tuple[str, Optional[str], np.ndarray]]:
        """Objects with a class name whose nearest boundary point lies within ``max_distance`` of ``point``:
        ``{object index: (class_name, object_name, nearest point)}`` (_aabb_world.py:166-186; first minimum on ties, like
        ``argmin``).  A handful of boxes on the host -- the SemanticDetector is not on the accelerated path."""
        from .util import closest_point_of_segment
        out = {}
        for idx, names in self._state.semantic_objects.items():
            near = closest_point_of_segment(point, self._state.segments[4 * idx:4 * idx + 4])
            dist = np.linalg.norm(point - near, axis=1)
            k = int(np.argmin(dist))
            if dist[k] <= max_distance:
                out[idx] = names + (near[k],)
        return out
