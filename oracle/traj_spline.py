"""ORACLE (test infrastructure, not product code): TrajectoryFollowEnv's spline reference trajectory.

Restates ``TrajectoryFollowEnv._sample_sinusoid_trajectory`` + the ``info["target"|"target_vel"|"target_acc"]``
look-ups (envs/trajectory_follow_env.py:176-218, 162-168, 232-250) on the engine's Philox draw layout, calling the
reference's own dependency -- ``scipy.interpolate.CubicSpline(bc_type='natural')`` (:211) -- for the spline itself.

Draw layout (uint32 words of Philox4x32-10, key = seed, counter = (global env id, episode, block, stream)):
    stream 0 (reset), block 0, words 0..2 : start position  ~ U(init_lo, init_hi)      (utils/state.py:90-98)
    stream 3 (traj),  block 0, words 0..2 : trajectory centre ~ U(center_lo, center_hi)  (:232)
                       block 0, word 3     : n_wp = 3 + min(floor(u * 3), 2)             (:194, integers(3, 6))
                       blocks 1..4, word 5*axis + j (j < 5): waypoint offset ~ U(-amp[axis], amp[axis])  (:203-205)
The reference draws ``n_wp`` offsets per axis from one PCG64 stream; here five are drawn per axis and the first
``n_wp`` are used, so that a draw's position does not depend on ``n_wp``.  The first waypoint is the start position
(:208-209).  Uniforms are float32 (fl32(fl32(u * (hi - lo)) + lo)); everything after the draws is float64 like the
reference, and the sampled arrays are cast to float32 (:197-199).
"""
from __future__ import annotations

import numpy as np
from scipy.interpolate import CubicSpline

from . import philox

STREAM_TRAJ = 3


def draws(cfg, env_ids, episode):
    """-> start[k,3] f32, centre[k,3] f32, n_wp[k] int, offsets[k,3,5] f32."""
    ids = np.asarray(env_ids, dtype=np.uint32)
    raw0 = philox.draw_blocks(cfg.seed, ids, episode, 1, philox.STREAM_RESET)
    start = np.stack([philox.uniform(raw0[:, i], cfg.init_lo[i], cfg.init_hi[i]) for i in range(3)], axis=1)
    raw = philox.draw_blocks(cfg.seed, ids, episode, 5, STREAM_TRAJ)
    centre = np.stack([philox.uniform(raw[:, i], cfg.traj_center_lo[i], cfg.traj_center_hi[i]) for i in range(3)], axis=1)
    n_wp = 3 + np.minimum((philox.u01(raw[:, 3]) * np.float32(3.0)).astype(np.int64), 2)
    off = np.zeros((len(ids), 3, 5), dtype=np.float32)
    for a in range(3):
        for j in range(5):
            off[:, a, j] = philox.uniform(raw[:, 4 + 5 * a + j], -cfg.traj_amp[a], cfg.traj_amp[a])
    return start.astype(np.float32), centre.astype(np.float32), n_wp, off


def waypoints(cfg, env_ids, episode):
    """-> list of (wp_times[n], wp_positions[n,3]) float64 per env."""
    start, centre, n_wp, off = draws(cfg, env_ids, episode)
    T = sample_times(cfg)[-1]
    out = []
    for k in range(len(n_wp)):
        n = int(n_wp[k])
        y = centre[k].astype(np.float64)[None, :] + off[k, :, :n].astype(np.float64).T
        y[0] = start[k].astype(np.float64)
        out.append((np.linspace(0.0, T, n), y))
    return out


def sample_times(cfg):
    """envs/trajectory_follow_env.py:187-191."""
    N = int(cfg.max_episode_steps)
    if cfg.spline_duration is not None:
        return np.linspace(0.0, float(cfg.spline_duration), N)
    return np.arange(N) * float(cfg.dt_nominal)


def info(cfg, env_ids, episode, idx):
    """target / target_vel / target_acc at sample index idx[k] of each env's spline -> float32 [k, 9] (scipy)."""
    idx = np.asarray(idx, dtype=np.int64)
    t = sample_times(cfg)
    out = np.zeros((len(idx), 9), dtype=np.float32)
    for k, (wt, wy) in enumerate(waypoints(cfg, env_ids, episode)):
        for a in range(3):
            cs = CubicSpline(wt, wy[:, a], bc_type="natural")
            tk = t[idx[k]]
            out[k, a] = np.float32(cs(tk))
            out[k, 3 + a] = np.float32(cs.derivative(1)(tk))
            out[k, 6 + a] = np.float32(cs.derivative(2)(tk))
    return out


def natural_spline_closed_form(y, T, t):
    """The device algorithm (csrc/qs_traj.cuh), float64: natural cubic spline through n in {3, 4, 5} equally spaced
    knots on [0, T], second-derivative form, Thomas elimination; returns (s, s', s'') at t."""
    y = np.asarray(y, dtype=np.float64)
    n = len(y)
    h = T / (n - 1)
    m = np.zeros(n)
    if n > 2:
        d = [6.0 * (y[i - 1] - 2.0 * y[i] + y[i + 1]) / (h * h) for i in range(1, n - 1)]
        k = n - 2
        c = [0.0] * k          # modified super-diagonal
        r = [0.0] * k
        c[0] = 1.0 / 4.0
        r[0] = d[0] / 4.0
        for i in range(1, k):
            den = 4.0 - c[i - 1]
            c[i] = 1.0 / den
            r[i] = (d[i] - r[i - 1]) / den
        x = [0.0] * k
        x[k - 1] = r[k - 1]
        for i in range(k - 2, -1, -1):
            x[i] = r[i] - c[i] * x[i + 1]
        m[1:n - 1] = x
    j = min(int(t / h), n - 2)
    a = (j + 1) * h - t
    b = t - j * h
    s = (m[j] * a ** 3 + m[j + 1] * b ** 3) / (6.0 * h) + (y[j] - m[j] * h * h / 6.0) * a / h + (y[j + 1] - m[j + 1] * h * h / 6.0) * b / h
    s1 = (-m[j] * a * a + m[j + 1] * b * b) / (2.0 * h) + (y[j + 1] - y[j]) / h - (m[j + 1] - m[j]) * h / 6.0
    s2 = (m[j] * a + m[j + 1] * b) / h
    return s, s1, s2
