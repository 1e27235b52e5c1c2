"""Host helpers of ``pyminisim.util`` (util/__init__.py:1-5): angle arithmetic and point / segment geometry.

Small host-side functions the drop-in classes and user scripts call; the device code has its own versions of the same
formulas (pms_common.cuh).  Results equal the reference's bit for bit (tests/test_eth_record.py, golden fixture).
"""
from __future__ import annotations

import numpy as np

_TWO_PI = 2 * np.pi


def wrap_angle(angle):
    """Angle(s) to [-pi, pi) by floor-mod (util/_util.py:5-6)."""
    return (angle + np.pi) % _TWO_PI - np.pi


def angle_linspace(start_angle: float, end_angle: float, n: int) -> np.ndarray:
    """n + 1 headings from ``start_angle`` in n equal increments of the wrapped difference, each brought back into
    [-pi, pi] by one turn when it leaves it (util/_util.py:9-23; note n increments, not n points -- the last item is the
    wrapped end angle up to rounding)."""
    inc = wrap_angle(end_angle - start_angle) / n
    out = np.empty(n + 1)
    out[0] = cur = start_angle
    for k in range(1, n + 1):
        cur = cur + inc
        if np.abs(cur) > np.pi:
            cur = cur - _TWO_PI if cur > 0.0 else cur + _TWO_PI
        out[k] = cur
    return out


def closest_point_of_segment(point: np.ndarray, segment: np.ndarray) -> np.ndarray:
    """Closest point(s) of segment(s) ``(x1, y1, x2, y2)`` to point(s): (2,) / (N, 2) against (4,) / (M, 4) ->
    (2,) / (N, 2) / (M, 2) / (N, M, 2) (util/_geometry.py:4-37)."""
    assert point.shape == (2,) or (len(point.shape) == 2 and point.shape[1] == 2), \
        f"Point must have shape (N, 2) or (2,), got {point.shape}"
    assert segment.shape == (4,) or (len(segment.shape) == 2 and segment.shape[1] == 4), \
        f"Segment must have shape (4,) or (M, 4), got {segment.shape}"
    one_point, one_segment = point.ndim == 1, segment.ndim == 1
    pts = point[None] if one_point else point
    seg = segment[None] if one_segment else segment
    a, ab = seg[:, :2], seg[:, 2:] - seg[:, :2]
    rel = pts[:, None, :] - a[None]
    t = np.clip(np.einsum("ijk,jk->ij", rel, ab) / np.einsum("ij,ij->i", ab, ab), 0, 1)
    out = a[None] + t[:, :, None] * ab[None]
    if one_segment:
        out = out[:, 0, :]
    if one_point:
        out = out[0]
    return out


def point_to_segment_distance(point: np.ndarray, segment: np.ndarray):
    """Distance(s) from point(s) to segment(s) (util/_geometry.py:40-43)."""
    return np.linalg.norm(point - closest_point_of_segment(point, segment), axis=-1)
