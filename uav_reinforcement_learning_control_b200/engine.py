"""ctypes binding of libquadsim.so (include/quadsim_abi.h) on torch CUDA tensors.

PyTorch is plumbing only here: it owns device memory and streams; every computation is a
hand-written sm_100a kernel reached through the C ABI.  There is no CPU path: a missing
library, a missing GPU or a non-zero status raises ``QuadSimError``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import config as Q
from . import model as M
from .build import LIB_PATH

__all__ = ["QuadSimError", "load_library", "Engine"]


class QuadSimError(RuntimeError):
    pass


_LIB = None

_EXPORTS = {
    "qs_abi_version": (C.c_int, []),
    "qs_last_error_string": (C.c_char_p, []),
    "qs_params_size": (C.c_int, []),
    "qs_launch_count": (C.c_uint64, []),
    "qs_create": (C.c_int, [C.POINTER(Q.QsParams), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "qs_destroy": (C.c_int, [C.c_void_p]),
    "qs_num_envs": (C.c_int, [C.c_void_p]),
    "qs_get_params": (C.c_int, [C.c_void_p, C.POINTER(Q.QsParams)]),
    "qs_reset": (C.c_int, [C.c_void_p] + [C.c_void_p] * 4 + [C.c_void_p]),
    "qs_step": (C.c_int, [C.c_void_p] + [C.c_void_p] * 9 + [C.c_void_p]),
    "qs_observe": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.c_void_p]),
    "qs_physics_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qs_rollout_random": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qs_policy_param_count": (C.c_int, [C.POINTER(Q.QsPolicyDesc)]),
    "qs_rollout_policy": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Q.QsPolicyDesc), C.c_void_p, C.c_int32, C.c_uint32]
                          + [C.c_void_p] * 9 + [C.c_void_p, C.c_void_p]),
    "qs_gae": (C.c_int, [C.c_int32, C.c_int32] + [C.c_void_p] * 5 + [C.c_float, C.c_float, C.c_int32]
               + [C.c_void_p] * 2 + [C.c_void_p]),
    "qs_traj_info": (C.c_int, [C.c_void_p] + [C.c_void_p] * 4 + [C.c_void_p]),
    "qs_step_host": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.c_void_p]),
    "qs_step_host_ex": (C.c_int, [C.c_void_p] + [C.c_void_p] * 6 + [C.c_void_p]),
    "qs_step_host_bytes": (C.c_int, [C.c_void_p] + [C.c_void_p] * 6 + [C.c_void_p]),
    "qs_ppo_workspace_bytes": (C.c_int64, [C.POINTER(Q.QsPolicyDesc)]),
    "qs_ppo_grad": (C.c_int, [C.POINTER(Q.QsPolicyDesc)] + [C.c_void_p] * 7
                    + [C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32] + [C.c_void_p] * 3),
    "qs_ppo_permutation": (C.c_int, [C.c_int32, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]),
    "qs_ppo_comm_create": (C.c_int, [C.POINTER(Q.QsPolicyDesc), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "qs_ppo_comm_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "qs_ppo_comm_import": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "qs_ppo_comm_slot": (C.c_void_p, [C.c_void_p, C.c_uint32]),
    "qs_ppo_adam_peer": (C.c_int, [C.POINTER(Q.QsPolicyDesc), C.c_void_p, C.c_uint32] + [C.c_void_p] * 3 + [C.c_int32]
                         + [C.c_float] * 5 + [C.c_void_p] * 3),     # (timeout: QS_PEER_TIMEOUT_MS in the environment)
    "qs_ppo_comm_error": (C.c_int, [C.c_void_p]),
    "qs_ppo_comm_close_peers": (C.c_int, [C.c_void_p]),
    "qs_ppo_comm_destroy": (C.c_int, [C.c_void_p]),
    "qs_ppo_pack": (C.c_int, [C.POINTER(Q.QsPolicyDesc)] + [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_void_p]),
    "qs_ppo_grad_packed": (C.c_int, [C.POINTER(Q.QsPolicyDesc)] + [C.c_void_p] * 4
                           + [C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32] + [C.c_void_p] * 3),
    "qs_obs_stats_workspace_bytes": (C.c_int64, [C.c_int32]),
    "qs_obs_stats_update": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                      C.c_void_p, C.c_void_p]),
    "qs_ppo_adam": (C.c_int, [C.POINTER(Q.QsPolicyDesc)] + [C.c_void_p] * 4 + [C.c_int32] + [C.c_float] * 6
                    + [C.c_void_p] * 2),
    "qs_ppo_update_epoch": (C.c_int, [C.POINTER(Q.QsPolicyDesc)] + [C.c_void_p] * 4 + [C.c_int32, C.c_int32]
                            + [C.c_float] * 3 + [C.c_int32] + [C.c_void_p] * 2 + [C.c_int32] + [C.c_float] * 5
                            + [C.c_void_p, C.c_uint32] + [C.c_void_p] * 4 + [C.c_uint64, C.c_void_p]),
}


def load_library(path: str | None = None):
    """dlopen libquadsim.so and type every export of include/quadsim_abi.h.  Loud on failure."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    path = path or os.environ.get("QS_LIB_PATH") or LIB_PATH     # QS_LIB_PATH: tuning variants for A/B runs
    if not os.path.exists(path):
        raise QuadSimError(f"{path} is missing: run `python -m uav_reinforcement_learning_control_b200.build` "
                           "(or __graft_entry__.build()).  There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in _EXPORTS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise QuadSimError(f"libquadsim.so does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if lib.qs_abi_version() != 1:
        raise QuadSimError("libquadsim ABI version mismatch")
    if lib.qs_params_size() != C.sizeof(Q.QsParams):
        raise QuadSimError(f"QsParams mirror out of sync: C {lib.qs_params_size()} B vs ctypes {C.sizeof(Q.QsParams)} B")
    _LIB = lib
    return lib


def exported_symbols():
    return list(_EXPORTS)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class Engine:
    """One handle = one shard of envs on one device (qs_create ... qs_destroy)."""

    def __init__(self, cfg: Q.EnvConfig, num_envs: int, device=0, xml_path: str | None = None,
                 assume_no_contact: bool = False):
        import torch
        if not torch.cuda.is_available():
            raise QuadSimError("no CUDA device: the quadsim engine has no CPU fallback")
        self.lib = load_library()
        self.torch = torch
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        tree = M.load_mjcf(xml_path or M.default_model_path())
        self.tree = tree
        self.constants = M.derive_constants(tree)
        # the kernels have no collision stage: refuse a model whose geoms could touch inside the termination bounds
        M.check_contacts(tree, *cfg.position_bounds(), overshoot=cfg.position_overshoot(self.constants.dt),
                         assume_no_contact=assume_no_contact)
        self.cfg = cfg
        self.num_envs = int(num_envs)
        self.obs_dim = cfg.obs_dim
        self.params = Q.pack_params(self.constants, cfg)
        table = cfg.target_table() if cfg.mode in (Q.MODE_MJX_BRAX, Q.MODE_MJX_PLAYGROUND) else None
        wps = cfg.waypoint_table() if cfg.waypoint_mode else None
        self._table, self._wps = table, wps
        h = C.c_void_p()
        rc = self.lib.qs_create(C.byref(self.params), self.num_envs, self.device.index,
                                None if table is None else table.ctypes.data_as(C.c_void_p),
                                None if wps is None else wps.ctypes.data_as(C.c_void_p), C.byref(h))
        self._check(rc)
        self.handle = h

    # ------------------------------------------------------------------ helpers
    def _check(self, rc):
        if rc != 0:
            raise QuadSimError(f"libquadsim error {rc}: {self.lib.qs_last_error_string().decode()}")

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _f32(self, *shape):
        return self.torch.empty(shape, dtype=self.torch.float32, device=self.device)

    def _chk(self, t, shape, name):
        torch = self.torch
        if t is None:
            return
        if (t.dtype != torch.float32 or not t.is_cuda or t.device != self.device or not t.is_contiguous()
                or tuple(t.shape) != tuple(shape)):
            raise QuadSimError(f"{name}: expected contiguous float32 tensor of shape {tuple(shape)} on {self.device}, "
                               f"got {t.dtype} {tuple(t.shape)} on {t.device}")

    def new_state(self):
        st = self.torch.zeros((Q.NPLANES, self.num_envs), dtype=self.torch.float32, device=self.device)
        st[3] = 1.0
        return st

    def close(self):
        if getattr(self, "handle", None) is not None:
            self.lib.qs_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ ABI calls
    def reset(self, state, mask=None, obs=None, first_state=None):
        n, D = self.num_envs, self.obs_dim
        self._chk(state, (Q.NPLANES, n), "state")
        if obs is None:
            obs = self._f32(n, D)
        self._chk(obs, (n, D), "obs"); self._chk(first_state, (21, n), "first_state")
        if mask is not None and (mask.dtype != self.torch.uint8 or tuple(mask.shape) != (n,) or mask.device != self.device):
            raise QuadSimError(f"mask must be a uint8 tensor of shape (num_envs,) on {self.device}")
        self._check(self.lib.qs_reset(self.handle, _ptr(state), _ptr(mask), _ptr(obs), _ptr(first_state), self._stream()))
        return obs

    def step(self, state, action, obs=None, reward=None, done=None, truncated=None, metrics=None,
             terminal_obs=None, first_state=None):
        n, D = self.num_envs, self.obs_dim
        self._chk(state, (Q.NPLANES, n), "state"); self._chk(action, (n, 4), "action")
        obs = self._f32(n, D) if obs is None else obs
        reward = self._f32(n) if reward is None else reward
        done = self._f32(n) if done is None else done
        self._chk(obs, (n, D), "obs"); self._chk(reward, (n,), "reward"); self._chk(done, (n,), "done")
        self._chk(truncated, (n,), "truncated"); self._chk(metrics, (4, n), "metrics")
        self._chk(terminal_obs, (n, D), "terminal_obs"); self._chk(first_state, (21, n), "first_state")
        self._check(self.lib.qs_step(self.handle, _ptr(state), _ptr(action), _ptr(obs), _ptr(reward), _ptr(done),
                                     _ptr(truncated), _ptr(metrics), _ptr(terminal_obs), _ptr(first_state),
                                     self._stream()))
        return obs, reward, done

    def observe(self, state, action=None):
        n, D = self.num_envs, self.obs_dim
        self._chk(state, (Q.NPLANES, n), "state"); self._chk(action, (n, 4), "action")
        obs, reward, done = self._f32(n, D), self._f32(n), self._f32(n)
        self._check(self.lib.qs_observe(self.handle, _ptr(state), _ptr(action), _ptr(obs), _ptr(reward), _ptr(done),
                                        self._stream()))
        return obs, reward, done

    def traj_info(self, state, episode=None, sample_index=None, out=None):
        """TrajectoryFollowEnv info target | target_vel | target_acc, [n, 9] (include/quadsim_abi.h: qs_traj_info)."""
        n = self.num_envs
        self._chk(state, (Q.NPLANES, n), "state")
        out = self._f32(n, 9) if out is None else out
        self._chk(out, (n, 9), "out")
        for name, t, dt in (("episode", episode, self.torch.int32), ("sample_index", sample_index, self.torch.int32)):
            if t is not None and (t.dtype != dt or tuple(t.shape) != (n,) or not t.is_contiguous() or t.device != out.device):
                raise QuadSimError(f"{name}: expected a contiguous int32 [{n}] tensor on {out.device}")
        self._check(self.lib.qs_traj_info(self.handle, _ptr(state), _ptr(episode), _ptr(sample_index), _ptr(out),
                                          self._stream()))
        return out

    def physics_step(self, state, ctrl):
        self._chk(state, (Q.NPLANES, self.num_envs), "state"); self._chk(ctrl, (self.num_envs, 4), "ctrl")
        self._check(self.lib.qs_physics_step(self.handle, _ptr(state), _ptr(ctrl), self._stream()))

    def rollout_random(self, state, T, t0=0, stats=None, first_state=None):
        n = self.num_envs
        self._chk(state, (Q.NPLANES, n), "state"); self._chk(stats, (4, n), "stats")
        self._chk(first_state, (21, n), "first_state")
        self._check(self.lib.qs_rollout_random(self.handle, _ptr(state), int(T), int(t0) & 0xFFFFFFFF, _ptr(stats),
                                               _ptr(first_state), self._stream()))

    def policy_desc(self, dist=0, deterministic=False, bootstrap_gamma=0.0, tensor_cores=False):
        d = Q.QsPolicyDesc()
        d.obs_dim, d.hidden, d.act_dim, d.dist = self.obs_dim, 128, 4, int(dist)
        d.deterministic, d.bootstrap_gamma = int(deterministic), float(bootstrap_gamma)
        d.tensor_cores = int(tensor_cores)
        return d

    def policy_param_count(self, dist=0):
        d = self.policy_desc(dist)
        return int(self.lib.qs_policy_param_count(C.byref(d)))

    def rollout_policy(self, state, params, T, t0=0, dist=0, deterministic=False, bootstrap_gamma=0.0, tensor_cores=False,
                       buffers=None, first_state=None):
        """Runs T fused policy+env steps.  `buffers`: dict of optional preallocated trajectory tensors."""
        torch = self.torch
        n, D = self.num_envs, self.obs_dim
        d = self.policy_desc(dist, deterministic, bootstrap_gamma, tensor_cores)
        self._chk(state, (Q.NPLANES, n), "state")
        self._chk(params, (self.policy_param_count(dist),), "policy params")
        b = dict(buffers or {})
        shapes = {"last_obs": (n, D), "obs": (T, n, D), "act": (T, n, 4), "logp": (T, n), "value": (T, n),
                  "reward": (T, n), "done": (T, n), "trunc": (T, n), "last_value": (n,)}
        for k, shp in shapes.items():
            if k not in b:
                b[k] = torch.empty(shp, dtype=torch.float32, device=self.device)
            elif b[k] is not None:
                self._chk(b[k], shp, k)
        self._check(self.lib.qs_rollout_policy(
            self.handle, _ptr(state), C.byref(d), _ptr(params), int(T), int(t0) & 0xFFFFFFFF, _ptr(b["last_obs"]),
            _ptr(b["obs"]), _ptr(b["act"]), _ptr(b["logp"]), _ptr(b["value"]), _ptr(b["reward"]), _ptr(b["done"]),
            _ptr(b["trunc"]), _ptr(b["last_value"]), _ptr(first_state), self._stream()))
        return b

    def gae(self, reward, value, done, trunc, last_value, gamma, lam, brax_form=False, adv=None, ret=None):
        torch = self.torch
        T, B = reward.shape
        for name, t in (("reward", reward), ("value", value), ("done", done)):
            self._chk(t, (T, B), name)
        self._chk(trunc, (T, B), "trunc"); self._chk(last_value, (B,), "last_value")
        adv = torch.empty_like(reward) if adv is None else adv
        ret = torch.empty_like(reward) if ret is None else ret
        self._check(self.lib.qs_gae(T, B, _ptr(reward), _ptr(value), _ptr(done), _ptr(trunc), _ptr(last_value),
                                    float(gamma), float(lam), int(brax_form), _ptr(adv), _ptr(ret), self._stream()))
        return adv, ret

    def step_host(self, state, action_host: np.ndarray, obs_host: np.ndarray, reward_host: np.ndarray,
                  done_host: np.ndarray, trunc_host: np.ndarray | None = None):
        """Env.step with page-locked HOST buffers: H2D action, step, D2H obs / reward / done (/ truncated), synchronise.
        Flags come back in the dtype of `done_host`: uint8 / bool (qs_step_host_bytes: what Gymnasium and SB3 return, and 3
        bytes per env-step less on the PCIe link) or float32 (qs_step_host_ex).  Pageable buffers are refused with QS_EINVAL."""
        u8 = done_host.dtype.itemsize == 1
        if trunc_host is not None and (trunc_host.dtype.itemsize == 1) != u8:
            raise QuadSimError("step_host: done_host and trunc_host must have the same dtype")
        if not u8 and done_host.dtype != np.float32:
            raise QuadSimError("step_host: flags must be uint8 / bool or float32")
        fn = self.lib.qs_step_host_bytes if u8 else self.lib.qs_step_host_ex
        self._check(fn(self.handle, _ptr(state), action_host.ctypes.data_as(C.c_void_p),
                                             obs_host.ctypes.data_as(C.c_void_p), reward_host.ctypes.data_as(C.c_void_p),
                                             done_host.ctypes.data_as(C.c_void_p),
                                             None if trunc_host is None else trunc_host.ctypes.data_as(C.c_void_p),
                                             self._stream()))

    def launch_count(self):
        return int(self.lib.qs_launch_count())
