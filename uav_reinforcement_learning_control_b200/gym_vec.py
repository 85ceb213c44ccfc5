"""Gymnasium ``VectorEnv`` / SB3 ``VecEnv`` facade of the hover env over the engine.

Mirrors ``HoverEnv`` (reference: envs/hover_env.py:13-238) as the vectorised env the reference
trainers build with ``make_vec_env(make_env, n_envs=16)`` (train.py:48): same spaces, bounds,
attribute names used by the reference's wrappers / evaluators (``dt``, ``frame_skip``,
``_prev_action``, ``target_state``, ``set_state``, ``_get_obs``, ``_obs_bounds`` ...), same
auto-reset contract (``terminal_observation`` / ``TimeLimit.truncated``).  gymnasium and SB3 are
not installed in this image, so the facade is duck-typed; ``Box`` below carries only what the
reference touches (low/high/shape/dtype/contains/sample).

Two call styles:
  * torch CUDA tensors in / out (zero copy, the fast path), and
  * NumPy in / out (what SB3 passes): routed through ``qs_step_host`` -- one H2D of the actions and
    one D2H of obs / reward / done per step, pinned staging buffers.
"""
from __future__ import annotations

import numpy as np

from . import config as Q
from .engine import Engine

__all__ = ["Box", "HoverVecEnv"]


class Box:
    """Minimal stand-in for ``gymnasium.spaces.Box`` (only what the reference uses)."""

    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        self.dtype = np.dtype(dtype)
        low = np.asarray(low, dtype=self.dtype); high = np.asarray(high, dtype=self.dtype)
        if shape is not None:
            low = np.broadcast_to(low, shape).copy(); high = np.broadcast_to(high, shape).copy()
        self.low, self.high, self.shape = low, high, low.shape
        self._rng = np.random.default_rng(seed)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)


class _TargetState:
    """``env.target_state.state[0:3]`` / ``.position`` as the reference's evaluators poke it (evaluate.py:497)."""

    def __init__(self, env):
        self._env = env

    @property
    def position(self):
        return self._env._planes[21:24].t()

    @property
    def state(self):
        return self.position


class HoverVecEnv:
    def __init__(self, num_envs: int = 16, device=0, *, max_episode_steps: int = 512, battery: bool = True,
                 auto_reset: bool = True, seed: int = 0, env_id_offset: int = 0, cfg: Q.EnvConfig | None = None,
                 xml_path: str | None = None, sb3_infos: bool | None = None, wrapper: str | None = None):
        import torch
        self.torch = torch
        # wrapper: the reference's WRAPPER_REGISTRY names (envs/wrappers.py:28-29).  "RateControlWrapper" (the SB3
        # trainer's default, train.py:31) is fused into the step kernel as an action pre-stage; "RelPosActWrapper"
        # (7-D obs = normalised rel-pos + previous action, envs/wrappers.py:13-25) is a view built here.
        if wrapper not in (None, "none", "RateControlWrapper", "RelPosActWrapper"):
            raise KeyError(wrapper)
        self.wrapper = None if wrapper in (None, "none") else wrapper
        if cfg is None:
            cfg = Q.EnvConfig.hover_gym(battery=battery, max_episode_steps=max_episode_steps, seed=seed,
                                        env_id_offset=env_id_offset, rate_wrapper=(wrapper == "RateControlWrapper"),
                                        auto_reset=Q.RESET_RESAMPLE if auto_reset else Q.RESET_NONE)
        self.cfg = cfg
        self.engine = Engine(cfg, num_envs, device=device, xml_path=xml_path)
        self.num_envs = int(num_envs)
        self.device = self.engine.device
        D = cfg.obs_dim
        # spaces / bounds (hover_env.py:31-65)
        self.single_action_space = Box(-1.0, 1.0, (4,), np.float32)
        self.single_observation_space = Box(-1.0, 1.0, (7 if self.wrapper == "RelPosActWrapper" else D,), np.float32)
        self.action_space = self.single_action_space
        self.observation_space = self.single_observation_space
        self._obs_bounds = Box(cfg.obs_lo, cfg.obs_hi)
        self._state_bounds = Box(cfg.term_lo, cfg.term_hi)
        self._initial_state_bounds = Box(cfg.init_lo, cfg.init_hi)
        self._target_pos_bounds = Box(cfg.target_lo, cfg.target_hi)
        self._action_bounds = Box(cfg.act_lo, cfg.act_hi)
        self.max_motor_thrust = cfg.max_motor_thrust
        self.max_total_thrust = 4 * cfg.max_motor_thrust
        self.max_torque = Q.MAX_TORQUE
        self.max_episode_steps = cfg.max_episode_steps
        self.A_inv = cfg.mixer()[1]
        self.dt = self.engine.constants.dt
        self.frame_skip = 1
        self.nominal_voltage, self.min_voltage = cfg.v_nominal, cfg.v_min
        self.render_mode = None
        self.target_state = _TargetState(self)
        f32 = dict(dtype=torch.float32, device=self.device)
        n = self.num_envs
        self._planes = self.engine.new_state()
        self._obs = torch.zeros(n, D, **f32); self._rew = torch.zeros(n, **f32)
        self._term = torch.zeros(n, **f32); self._trunc = torch.zeros(n, **f32)
        self._terminal_obs = torch.zeros(n, D, **f32)
        self._prev_action = torch.zeros(n, 4, **f32)
        self._pending = None
        self._sb3_infos = (n <= 4096) if sb3_infos is None else sb3_infos
        self._host = None
        self.reset()

    # ------------------------------------------------------------------ helpers
    @property
    def unwrapped(self):
        return self

    def _episode_seed(self, seed):
        if seed is not None:
            # re-key the Philox stream: episode counters restart from a seed-derived word
            w = (int(seed) * 2654435761) & 0x7FFFFFFF
            self._planes[26] = self.torch.full((self.num_envs,), w, dtype=self.torch.int32, device=self.device).view(self.torch.float32)

    def state12(self):
        """info["state"]: float32 [B, 12] = pos, rpy, v, w (utils/state.py:28-46), recovered from obs."""
        torch = self.torch
        lo = torch.tensor(self._obs_bounds.low, device=self.device); hi = torch.tensor(self._obs_bounds.high, device=self.device)
        s = (self._obs + 1.0) * 0.5 * (hi - lo) + lo
        s[:, 0:3] = self._planes[0:3].t()
        return s

    def _wrap_obs(self, obs):
        if self.wrapper == "RelPosActWrapper":
            return self.torch.cat([obs[:, 0:3], self._prev_action], dim=1)
        return obs

    @property
    def max_rate_rad(self):
        return float(np.deg2rad(self.cfg.rate_max_deg))

    @property
    def _rate_int_torque(self):
        return self._planes[32:35].t()

    def _infos(self, finished=None, traj=None):
        info = {"target": self._planes[21:24].t(), "voltage": self._planes[25],
                "voltage_scale": (self._planes[25] / self.cfg.v_nominal).clamp(0.0, 1.0)}
        if traj is not None:
            # TrajectoryFollowEnv: the spline reference of the step that ran (trajectory_follow_env.py:162-168,245-250)
            info["target"], info["target_vel"], info["target_acc"] = traj[:, 0:3], traj[:, 3:6], traj[:, 6:9]
        if finished is not None:
            info["terminal_observation"] = self._terminal_obs
            info["final_observation"] = self._terminal_obs
            info["_final_observation"] = finished
            info["TimeLimit.truncated"] = (self._trunc != 0) & (self._term == 0)
        return info

    # ------------------------------------------------------------------ Gymnasium VectorEnv
    def reset(self, seed=None, options=None):
        self._episode_seed(seed)
        self.engine.reset(self._planes, obs=self._obs)
        self._prev_action = self.torch.zeros_like(self._prev_action)
        traj = self.engine.traj_info(self._planes) if self.cfg.mode == Q.MODE_TRAJ_GYM else None
        return self._wrap_obs(self._obs), self._infos(traj=traj)

    def step(self, actions):
        """-> (obs, reward, terminated, truncated, infos); torch in -> torch out, NumPy in -> NumPy out."""
        if isinstance(actions, np.ndarray):
            return self._step_numpy(actions)
        torch = self.torch
        a = torch.as_tensor(actions, dtype=torch.float32, device=self.device).contiguous()
        self._prev_action = a
        traj = None
        if self.cfg.mode == Q.MODE_TRAJ_GYM:
            # info index = step_count - 1 of the episode the step belongs to, also for envs the step auto-resets
            prev_sc = self._planes[24].view(torch.int32).clone(); prev_ep = self._planes[26].view(torch.int32).clone()
        self.engine.step(self._planes, a, obs=self._obs, reward=self._rew, done=self._term, truncated=self._trunc,
                         terminal_obs=self._terminal_obs)
        if self.cfg.mode == Q.MODE_TRAJ_GYM:
            traj = self.engine.traj_info(self._planes, episode=prev_ep, sample_index=prev_sc)
        finished = (self._term != 0) | (self._trunc != 0)
        if self.wrapper == "RelPosActWrapper" and bool(finished.any()):
            self._prev_action = self.torch.where(finished[:, None], self.torch.zeros_like(a), a)   # reset() zeroes it
        return self._wrap_obs(self._obs), self._rew, self._term != 0, self._trunc != 0, self._infos(finished, traj)

    def _step_numpy(self, actions):
        torch = self.torch
        n, D = self.num_envs, self.cfg.obs_dim
        if self._host is None:
            pin = lambda *s: torch.empty(s, dtype=torch.float32).pin_memory().numpy()
            pin8 = lambda *s: torch.empty(s, dtype=torch.uint8).pin_memory().numpy()
            # flags as bytes (qs_step_host_bytes) when the batch allows it: they ARE the bool arrays Gymnasium returns
            flag = pin8 if n % 4 == 0 else pin
            self._host = dict(act=pin(n, 4), obs=pin(n, D), rew=pin(n), done=flag(n), trunc=flag(n))
        h = self._host
        np.copyto(h["act"], np.asarray(actions, dtype=np.float32).reshape(n, 4))
        # terminated AND truncated come from the kernel (an env can be both; a waypoint lap is neither)
        self.engine.step_host(self._planes, h["act"], h["obs"], h["rew"], h["done"], h["trunc"])
        as_bool = lambda f: f.view(np.bool_).copy() if f.dtype == np.uint8 else f != 0
        return h["obs"].copy(), h["rew"].copy(), as_bool(h["done"]), as_bool(h["trunc"]), {}

    def close(self):
        self.engine.close()

    # ------------------------------------------------------------------ SB3 VecEnv facade
    def step_async(self, actions):
        self._pending = actions

    def step_wait(self):
        """-> (obs, rewards, dones, infos) with SB3's auto-reset contract (numpy)."""
        obs, rew, term, trunc, _ = self.step(self.torch.as_tensor(np.asarray(self._pending), device=self.device))
        dones = (term | trunc)
        d = dones.cpu().numpy()
        infos = []
        if self._sb3_infos:
            tobs = self._terminal_obs.cpu().numpy(); tr = (trunc & ~term).cpu().numpy()
            infos = [{"terminal_observation": tobs[i], "TimeLimit.truncated": bool(tr[i])} if d[i] else {}
                     for i in range(self.num_envs)]
        return obs.cpu().numpy(), rew.cpu().numpy(), d, infos

    def env_method(self, name, *args, **kw):
        return [getattr(self, name)(*args, **kw)]

    def get_attr(self, name, indices=None):
        return [getattr(self, name)] * self.num_envs

    # ------------------------------------------------------------------ HoverEnv single-env surface (batched)
    def set_state(self, qpos, qvel, indices=None):
        """hover_env.py:143-148: overwrite qpos[:7] / qvel[:6] (rotor state untouched)."""
        torch = self.torch
        qpos = torch.as_tensor(qpos, dtype=torch.float32, device=self.device).reshape(-1, 7)
        qvel = torch.as_tensor(qvel, dtype=torch.float32, device=self.device).reshape(-1, 6)
        idx = slice(None) if indices is None else indices
        self._planes[0:7, idx] = qpos.t().expand(7, self._planes[0:7, idx].shape[1]) if qpos.shape[0] == 1 else qpos.t()
        self._planes[11:17, idx] = qvel.t().expand(6, self._planes[11:17, idx].shape[1]) if qvel.shape[0] == 1 else qvel.t()

    def _get_obs(self):
        """hover_env.py:126-136 on the current state (no step)."""
        obs, _, _ = self.engine.observe(self._planes)
        self._obs = obs
        return obs
