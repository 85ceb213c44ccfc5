"""Waypoint tables for the waypoint-tracking mode (reference: utils/trajectories.py:6-81).

Same three closed curves at altitude z = centre_z, consecutive waypoints ~`spacing` apart,
re-derived here in vectorised form; tests/test_golden.py pins the tables against the
reference generator's output (tests/golden/reference_utils.json).
Returned as float64 [n, 3] arrays (the engine keeps them in float64: evaluate.py:540-541
measures the reach distance in double precision).
"""
from __future__ import annotations

import numpy as np

_DEFAULT_CENTER = (0.0, 0.0, 1.0)


def _center(center):
    return np.asarray(_DEFAULT_CENTER if center is None else center, dtype=np.float64)


def circle(spacing: float = 0.5, radius: float = 1.0, center=None) -> np.ndarray:
    """n = max(ceil(2 pi r / spacing), 4) points at equal angles, starting at angle 0."""
    c = _center(center)
    n = max(int(np.ceil(2.0 * np.pi * radius / spacing)), 4)
    ang = 2.0 * np.pi * np.arange(n) / n
    return np.stack([c[0] + radius * np.cos(ang), c[1] + radius * np.sin(ang), np.full(n, c[2])], axis=1)


def figure_eight(spacing: float = 0.5, radius: float = 1.0, center=None) -> np.ndarray:
    """Lemniscate x = r cos t, y = (r/2) sin 2t resampled at equal arc length.

    The arc length is the rectangle-rule cumulative sum over 1000 parameter samples and the
    parameter is recovered by linear interpolation of that table, which is the reference's
    discretisation (the waypoints depend on it at the 1e-3 level); n = max(ceil(L/spacing), 8).
    """
    c = _center(center)
    m = 1000
    t = np.linspace(0.0, 2.0 * np.pi, m, endpoint=False)
    speed = np.hypot(-radius * np.sin(t), radius * np.cos(2.0 * t))
    s = np.cumsum(speed * (2.0 * np.pi / m))
    n = max(int(np.ceil(s[-1] / spacing)), 8)
    tt = np.interp(np.linspace(0.0, s[-1], n, endpoint=False), s, t)
    return np.stack([c[0] + radius * np.cos(tt), c[1] + (radius / 2.0) * np.sin(2.0 * tt), np.full(n, c[2])], axis=1)


def square(spacing: float = 0.5, side_length: float = 1.5, center=None) -> np.ndarray:
    """Counter-clockwise from the (+,+) corner; each edge split into max(ceil(side/spacing), 1) segments."""
    c = _center(center)
    h = side_length / 2.0
    corners = np.array([[h, h], [-h, h], [-h, -h], [h, -h]], dtype=np.float64) + c[:2]
    pts = []
    for i in range(4):
        a, b = corners[i], corners[(i + 1) % 4]
        k = max(int(np.ceil(np.linalg.norm(b - a) / spacing)), 1)
        f = (np.arange(k) / k)[:, None]
        pts.append(a + f * (b - a))
    xy = np.concatenate(pts, axis=0)
    return np.concatenate([xy, np.full((len(xy), 1), c[2])], axis=1)


GENERATORS = {"eight": figure_eight, "circle": circle, "square": square}


def default_tables(spacing: float = 0.5):
    """The three shapes of BASELINE.json configs[3], in the order eight, circle, square."""
    return [GENERATORS[k](spacing=spacing) for k in ("eight", "circle", "square")]
