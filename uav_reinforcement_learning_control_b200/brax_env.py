"""Brax-protocol adapter: ``reset(rng) -> State``, ``step(state, action) -> State`` over the engine.

Mirrors the surface of the reference's Brax envs so that code written against them reads the
same (reference: train_brax_ppo.py:179-368 ``JaxMJXQuadBraxEnv``, :39-176 ``QuadHoverBraxEnv``;
consumers: ``ppo_train.train`` :589, evaluate_brax_ppo.py:106-125,300-301, test_brax_mixing.py:30-64):

    env.observation_size / action_size / backend / unwrapped
    env._ctrl_min / _ctrl_max / A_inv / max_motor_thrust / max_total_thrust / max_torque
    env._mix_to_motors(thrust, tau_x, tau_y, tau_z)
    state.pipeline_state.qpos / .qvel (.q / .qd), state.obs / reward / done / metrics / info
    info["time_out"], info["step_count"], info["traj_pos"]; wrapped: info["steps"], ["truncation"],
    ["first_pipeline_state"], ["first_obs"]

Differences that are inherent to the B200 design and documented in INTEGRATION.md: the env is
natively BATCHED (leading axis = num_envs, what brax's VmapWrapper would add), tensors are torch
CUDA tensors (every leaf supports ``__dlpack__`` for the JAX hand-off), and randomness is the
engine's Philox stream: ``reset(rng)`` uses rng only to pick the episode counter of each env.
With ``wrapped=True`` the Episode/AutoReset wrapper logic of ``brax.envs.training.wrap`` is fused
into the step kernel.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace as _dc_replace
from typing import Any, Dict

import numpy as np

from . import config as Q
from .engine import Engine

__all__ = ["State", "PipelineState", "QuadBraxEnv", "JaxMJXQuadBraxEnv", "QuadHoverBraxEnv"]


@dataclass
class PipelineState:
    """The fields of ``mjx.Data`` / brax pipeline state that downstream code reads."""
    planes: Any                    # float32 [32, B] engine state (include/quadsim_abi.h)
    ctrl: Any = None

    @property
    def qpos(self):
        return self.planes[0:11].t()

    @property
    def qvel(self):
        return self.planes[11:21].t()

    q = qpos
    qd = qvel

    def replace(self, **kw):
        return _dc_replace(self, **kw)


@dataclass
class State:
    """Same field names as ``brax.envs.base.State`` (train_brax_ppo.py:129,289)."""
    pipeline_state: PipelineState
    obs: Any
    reward: Any
    done: Any
    metrics: Dict[str, Any] = field(default_factory=dict)
    info: Dict[str, Any] = field(default_factory=dict)

    def replace(self, **kw):
        return _dc_replace(self, **kw)

    def tree_leaves(self):
        out = [self.pipeline_state.planes, self.obs, self.reward, self.done]
        out += list(self.metrics.values())
        out += [v for v in self.info.values() if hasattr(v, "__dlpack__")]
        return out

    def to_dlpack(self):
        """DLPack capsules of every tensor leaf (``jax.dlpack.from_dlpack`` on the consumer side)."""
        import torch.utils.dlpack as dl
        return {
            "planes": dl.to_dlpack(self.pipeline_state.planes), "obs": dl.to_dlpack(self.obs),
            "reward": dl.to_dlpack(self.reward), "done": dl.to_dlpack(self.done),
            "metrics": {k: dl.to_dlpack(v) for k, v in self.metrics.items()},
        }


class QuadBraxEnv:
    """Batched functional env with the reference's Brax-env attribute surface."""

    def __init__(self, xml_path: str | None = None, *, variant: str = "mjx", num_envs: int = 1024, device=0,
                 max_episode_steps: int = 500, traj_duration_seconds: float = 5.0,
                 pos_limit_xy: float = 3.0, pos_limit_z_low: float = 0.02, pos_limit_z_high: float = 4.0,
                 vel_limit: float = 20.0, target_height: float = 1.0, wrapped: bool = False,
                 episode_length: int = 500, seed: int = 0, env_id_offset: int = 0, impl: str = "b200",
                 backend: str = "mjx", n_frames: int = 1, action_min: float = 0.0, action_max: float = 13.0):
        import torch
        self.torch = torch
        if n_frames != 1:
            raise ValueError("n_frames != 1 is not supported (the reference uses 1: train_brax_ppo.py:44)")
        common = dict(pos_limit_xy=pos_limit_xy, z_low=pos_limit_z_low, z_high=pos_limit_z_high, seed=seed,
                      env_id_offset=env_id_offset,
                      episode_length=episode_length if wrapped else 0,
                      auto_reset=Q.RESET_RESTORE_FIRST if wrapped else Q.RESET_NONE)
        if variant == "mjx":
            cfg = Q.EnvConfig.mjx_brax(max_episode_steps=int(max_episode_steps), vel_limit=vel_limit,
                                       traj_duration_seconds=float(traj_duration_seconds), **common)
        elif variant == "hover":
            cfg = Q.EnvConfig.hover_brax(fixed_target=(0.0, 0.0, float(target_height)), **common)
        else:
            raise ValueError("variant must be 'mjx' or 'hover'")
        self.cfg = cfg
        self.engine = Engine(cfg, num_envs, device=device, xml_path=xml_path)
        self.num_envs = int(num_envs)
        self.device = self.engine.device
        self._backend = backend
        self._xml_path = xml_path
        self._max_episode_steps = int(max_episode_steps)
        self._traj_duration_seconds = float(traj_duration_seconds)
        self._pos_limit_xy, self._pos_limit_z_low, self._pos_limit_z_high = pos_limit_xy, pos_limit_z_low, pos_limit_z_high
        self._vel_limit = vel_limit
        # reference attribute surface (train_brax_ppo.py:211-230)
        self.max_motor_thrust = Q.MAX_MOTOR_THRUST
        self.max_total_thrust = 4 * self.max_motor_thrust
        self.max_torque = Q.MAX_TORQUE
        A, A_inv = cfg.mixer()
        f32 = dict(dtype=torch.float32, device=self.device)
        self.A_inv = torch.tensor(A_inv, **f32)
        self._ctrl_min = torch.tensor(cfg.act_lo, **f32)
        self._ctrl_max = torch.tensor(cfg.act_hi, **f32)
        self._target_pos = torch.tensor(cfg.fixed_target, **f32)
        self._traj_pos = torch.from_numpy(cfg.target_table()).to(self.device) if variant == "mjx" else None
        self._wrapped = bool(wrapped)
        self._first = None

    # ------------------------------------------------------------------ protocol
    @property
    def observation_size(self):
        return self.engine.obs_dim

    @property
    def action_size(self):
        return 4

    @property
    def backend(self):
        return self._backend

    @property
    def unwrapped(self):
        return self

    def _mix_to_motors(self, thrust, tau_x, tau_y, tau_z):
        """train_brax_ppo.py:291-305: F = clip(A_inv @ [T, tx, ty, tz], 0, max_motor_thrust)."""
        torch = self.torch
        u = torch.stack([torch.as_tensor(v, dtype=torch.float32, device=self.device)
                         for v in (thrust, tau_x, tau_y, tau_z)], dim=-1)
        return torch.clamp(u @ self.A_inv.T, 0.0, self.max_motor_thrust)

    def _episode_words(self, rng):
        """Map a Brax-style key ([2] or [B, 2] uint32, or an int) to one Philox episode word per env."""
        torch = self.torch
        if rng is None:
            return torch.zeros(self.num_envs, dtype=torch.int64, device=self.device)
        if isinstance(rng, int):
            return torch.full((self.num_envs,), rng & 0xFFFFFFFF, dtype=torch.int64, device=self.device)
        k = torch.as_tensor(np.asarray(rng.cpu() if hasattr(rng, "cpu") else rng).astype(np.int64), device=self.device)
        if k.ndim == 1:
            k = k.unsqueeze(0).expand(self.num_envs, -1)
        return (k[:, 0] * 2654435761 + k[:, -1]) & 0xFFFFFFFF

    def reset(self, rng=None) -> State:
        """train_brax_ppo.py:244-289 (batched).  rng selects the Philox episode counter of every env."""
        torch = self.torch
        eng = self.engine
        planes = eng.new_state()
        words = self._episode_words(rng)
        # uint32 word -> same bits as int32 -> stored in the float32 plane 26 (episode index)
        w32 = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
        planes[26] = w32.view(torch.float32)
        first = torch.empty(21, self.num_envs, dtype=torch.float32, device=self.device)
        obs = eng.reset(planes, first_state=first)
        zeros = torch.zeros(self.num_envs, dtype=torch.float32, device=self.device)
        info = {"time_out": zeros.clone(), "step_count": torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)}
        if self._traj_pos is not None:
            info["traj_pos"] = self._traj_pos            # shared [N, 3] table (the reference stores a copy per env)
        if self._wrapped:
            info.update(steps=zeros.clone(), truncation=zeros.clone(), first_pipeline_state=first, first_obs=obs.clone())
        metrics = {k: zeros.clone() for k in ("pos_error", "reward_hover", "reward_action", "reward")}
        return State(PipelineState(planes), obs, zeros.clone(), zeros.clone(), metrics, info)

    def step(self, state: State, action, donate: bool = False) -> State:
        """train_brax_ppo.py:307-356 (+ Episode/AutoReset wrappers when wrapped).  Functional: unless
        ``donate`` is true the input state's buffers are left untouched (one 128 B/env copy)."""
        torch = self.torch
        eng = self.engine
        n = self.num_envs
        planes = state.pipeline_state.planes if donate else state.pipeline_state.planes.clone()
        action = torch.as_tensor(action, dtype=torch.float32, device=self.device).contiguous()
        trunc = torch.empty(n, dtype=torch.float32, device=self.device)
        met = torch.empty(4, n, dtype=torch.float32, device=self.device)
        first = state.info.get("first_pipeline_state") if self._wrapped else None
        obs, rew, done = eng.step(planes, action, truncated=trunc, metrics=met, first_state=first)
        info = dict(state.info)
        info["time_out"] = done - trunc                   # Q3: the env's own failure flag (train_brax_ppo.py:353)
        info["step_count"] = planes[24].view(torch.int32)
        if self._wrapped:
            info["steps"] = planes[27].view(torch.int32).to(torch.float32)
            info["truncation"] = trunc
        metrics = {"pos_error": met[0], "reward_hover": met[1], "reward_action": met[2], "reward": met[3]}
        return State(PipelineState(planes), obs, rew, done, metrics, info)


def JaxMJXQuadBraxEnv(xml_path=None, **kw):
    """Drop-in constructor name for train_brax_ppo.py:179 (batched, B200-native)."""
    kw.pop("impl", None)
    return QuadBraxEnv(xml_path, variant="mjx", **kw)


def QuadHoverBraxEnv(xml_path=None, **kw):
    """Drop-in constructor name for train_brax_ppo.py:39."""
    return QuadBraxEnv(xml_path, variant="hover", **kw)
