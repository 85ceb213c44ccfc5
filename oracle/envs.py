"""ORACLE (test infrastructure, not product code) -- env semantics, PARITY UNPINNED for the physics.

NumPy restatement of the reference's five env variants around the physics step
(oracle/mujoco_pipeline.py), batched over the leading axis:

    mjx_brax        JaxMJXQuadBraxEnv     train_brax_ppo.py:179-368
    hover_gym       HoverEnv              envs/hover_env.py:13-238
    traj_gym        TrajectoryFollowEnv   envs/trajectory_follow_env.py:14-253
    hover_brax      QuadHoverBraxEnv      train_brax_ppo.py:39-176
    mjx_playground  JaxMJXQuadEnv         envs/jax_mjx_quad_env.py:40-183

plus the vectorising wrappers the trainers put around them (Brax EpisodeWrapper /
AutoResetWrapper [third party, restated from their published source]; SB3 VecEnv
auto-reset), the waypoint-advance rule of evaluate.py:535-557, and the engine's Philox
reset draws (oracle/philox.py).

Dtypes follow the reference: the physics state is float64 (C MuJoCo); everything the
reference casts to float32 before comparing (utils/state.py:25; all of the JAX envs) is
cast to float32 here before the same comparison, so masks / counters / indices are
decided on the same values the reference would decide them on.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.transform import Rotation

from uav_reinforcement_learning_control_b200 import config as Q

from . import philox
from .mujoco_pipeline import TreePipeline

F32 = np.float32


def _gym(mode):
    return mode in (Q.MODE_HOVER_GYM, Q.MODE_TRAJ_GYM)


def _brax(mode):
    return mode in (Q.MODE_MJX_BRAX, Q.MODE_HOVER_BRAX)


class OracleEnv:
    def __init__(self, tree, cfg: Q.EnvConfig):
        self.cfg = cfg
        self.pipe = TreePipeline(tree)
        self.A, self.A_inv = cfg.mixer()
        self.table = cfg.target_table() if cfg.mode in (Q.MODE_MJX_BRAX, Q.MODE_MJX_PLAYGROUND) else None
        self.obs_lo = np.asarray(cfg.obs_lo, dtype=F32); self.obs_hi = np.asarray(cfg.obs_hi, dtype=F32)
        self.term_lo = np.asarray(cfg.term_lo, dtype=F32); self.term_hi = np.asarray(cfg.term_hi, dtype=F32)
        self.act_lo = np.asarray(cfg.act_lo, dtype=F32); self.act_hi = np.asarray(cfg.act_hi, dtype=F32)

    # ------------------------------------------------------------------ state container
    @staticmethod
    def blank(n):
        qpos = np.zeros((n, 11)); qpos[:, 3] = 1.0
        return dict(qpos=qpos, qvel=np.zeros((n, 10)), target=np.zeros((n, 3), F32),
                    step_count=np.zeros(n, np.int32), voltage=np.zeros(n), episode=np.zeros(n, np.uint32),
                    ep_steps=np.zeros(n, np.int32), wp_idx=np.zeros(n, np.int32),
                    wp_reached=np.zeros(n, np.int32), laps=np.zeros(n, np.int32), done_prev=np.zeros(n, F32),
                    rate_int=np.zeros((n, 3)), prev_action=np.zeros((n, 4), F32))

    @staticmethod
    def from_planes(st):
        """float32 [32, n] engine state -> oracle state (float64 physics)."""
        n = st.shape[1]
        s = OracleEnv.blank(n)
        s["qpos"] = st[0:11].T.astype(np.float64); s["qvel"] = st[11:21].T.astype(np.float64)
        s["target"] = st[21:24].T.astype(F32).copy()
        s["step_count"] = st[24].view(np.int32).copy(); s["voltage"] = st[25].astype(np.float64)
        s["episode"] = st[26].view(np.uint32).copy(); s["ep_steps"] = st[27].view(np.int32).copy()
        s["wp_idx"] = st[28].view(np.int32).copy(); s["wp_reached"] = st[29].view(np.int32).copy()
        s["laps"] = st[30].view(np.int32).copy(); s["done_prev"] = st[31].copy()
        if st.shape[0] > 38:
            s["rate_int"] = st[32:35].T.astype(np.float64); s["prev_action"] = st[35:39].T.astype(F32).copy()
        return s

    # ------------------------------------------------------------------ pieces of step()
    def action_to_ctrl(self, action, voltage):
        """hover_env.py:169-177 / train_brax_ppo.py:309-314.  Returns (ctrl, new voltage)."""
        cfg = self.cfg
        a = np.asarray(action)
        u = (a.astype(np.float64) + 1.0) / 2.0 * (self.act_hi - self.act_lo).astype(np.float64) + self.act_lo
        if cfg.pre_clip_action:
            u = np.clip(u, self.act_lo, self.act_hi)
        F = np.clip(u @ self.A_inv.T, 0.0, cfg.max_motor_thrust)
        if cfg.battery:
            scale = np.clip(voltage / cfg.v_nominal, 0.0, 1.0)[:, None]
            F = np.clip(F * scale, 0.0, cfg.max_motor_thrust * scale)
            load = F.mean(axis=1) / max(cfg.max_motor_thrust, 1e-6)
            dV = (cfg.v_drop_base + cfg.v_drop_load * load) * self.pipe.dt
            voltage = np.clip(voltage - dV, cfg.v_min, cfg.v_nominal)
        return F, voltage

    def rate_to_torque(self, s, action):
        """RateControlWrapper.action (envs/rate_wrapper.py:69-98), float64 like the reference."""
        cfg = self.cfg
        a = np.asarray(action, dtype=F32)
        des = a[:, 1:4].astype(np.float64) * np.deg2rad(cfg.rate_max_deg)
        actual = s["qvel"][:, 3:6].astype(F32).astype(np.float64)      # QuadState keeps float32 (utils/state.py:25)
        err = des - actual
        tau_p = np.asarray(cfg.rate_inertia) * np.asarray(cfg.rate_kd) * err
        s["rate_int"] = np.clip(s["rate_int"] + cfg.rate_ki * self.pipe.dt * err, -cfg.rate_imax, cfg.rate_imax)
        tau_norm = np.clip((tau_p + s["rate_int"]) / Q.MAX_TORQUE, -1.0, 1.0)
        return np.concatenate([a[:, 0:1].astype(np.float64), tau_norm], axis=1).astype(F32)

    @staticmethod
    def state12(qpos, qvel):
        """utils/state.py:28-46: float32 [pos, rpy (scipy 'xyz'), v, w]."""
        n = qpos.shape[0]
        s = np.zeros((n, 12), F32)
        s[:, 0:3] = qpos[:, 0:3]
        quat = qpos[:, 3:7]
        ok = np.isfinite(quat).all(axis=1) & (np.linalg.norm(quat, axis=1) > 0)
        rpy = np.full((n, 3), np.nan)
        if ok.any():
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                rpy[ok] = Rotation.from_quat(quat[ok][:, [1, 2, 3, 0]]).as_euler("xyz")
        s[:, 3:6] = rpy
        s[:, 6:9] = qvel[:, 0:3]
        s[:, 9:12] = qvel[:, 3:6]
        return s

    def obs_gym(self, s12, target):
        """hover_env.py:126-136, utils/normalization.py:17 (float32 arithmetic as NumPy does it)."""
        x = s12.copy()
        x[:, 0:3] = target.astype(F32) - s12[:, 0:3]
        return (F32(2.0) * (x - self.obs_lo) / (self.obs_hi - self.obs_lo) - F32(1.0)).astype(F32)

    def evaluate(self, s, action):
        """Reward / done / obs of the state `s` (whose step_count is already advanced)."""
        cfg = self.cfg
        n = s["qpos"].shape[0]
        out = {}
        qpos32 = s["qpos"].astype(F32); qvel32 = s["qvel"].astype(F32)
        pos = qpos32[:, 0:3]
        if self.table is not None:
            idx = np.minimum(s["step_count"], cfg.max_episode_steps - 1)
            target = self.table[idx]
        elif cfg.mode == Q.MODE_HOVER_BRAX:
            target = np.broadcast_to(np.asarray(cfg.fixed_target, F32), (n, 3))
        else:
            target = s["target"].astype(F32)
        a = None if action is None else np.asarray(action, dtype=F32)
        asq = np.zeros(n, F32) if a is None else np.sum(a * a, axis=1, dtype=F32)
        out["truncated"] = np.zeros(n, F32)
        with np.errstate(all="ignore"):
            d = pos - target
            e_raw = np.sqrt(np.sum(d * d, axis=1, dtype=F32))
            if _gym(cfg.mode):
                s12 = self.state12(s["qpos"], s["qvel"])
                out["obs"] = self.obs_gym(s12, target)
                pe = e_raw.astype(np.float64)                      # float(np.linalg.norm(float32))
                out["pos_error"] = pe
                out["reward"] = np.exp(-pe ** 2)                   # hover_env.py:141
                out["reward_hover"] = out["reward"]; out["reward_action"] = np.zeros(n)
                inside = np.isfinite(s12).all(axis=1) & (s12 >= self.term_lo).all(axis=1) & (s12 <= self.term_hi).all(axis=1)
                out["done"] = (~inside).astype(F32)                # hover_env.py:150-157
                out["truncated"] = (s["step_count"] >= cfg.max_episode_steps).astype(F32)   # :188
                out["state12"] = s12
            else:
                obs = np.concatenate([qpos32, qvel32], axis=1)
                if cfg.mode == Q.MODE_MJX_BRAX:
                    fin = np.isfinite(qpos32).all(axis=1) & np.isfinite(qvel32).all(axis=1)
                    out_xy = (np.abs(pos[:, 0]) > cfg.pos_limit_xy) | (np.abs(pos[:, 1]) > cfg.pos_limit_xy)
                    out_z = (pos[:, 2] < F32(cfg.z_low)) | (pos[:, 2] > F32(cfg.z_high))
                    out_v = (np.abs(qvel32[:, 0:3]) > F32(cfg.vel_limit)).any(axis=1)
                    valid = fin & ~out_xy & ~out_z & ~out_v
                    pe = np.where(valid & np.isfinite(e_raw), e_raw, F32(1e3)).astype(F32)
                    rh = np.exp(-(pe.astype(np.float64) ** 2))
                    ra = -cfg.action_penalty * asq.astype(np.float64)
                    raw = rh + ra
                    out["pos_error"] = pe; out["reward_hover"] = rh; out["reward_action"] = ra
                    out["reward"] = np.where(valid & np.isfinite(raw), raw, -1.0)
                    out["done"] = np.where(valid, 0.0, 1.0).astype(F32)
                    obs = np.where(np.isfinite(obs), obs, F32(0.0))
                elif cfg.mode == Q.MODE_HOVER_BRAX:
                    pe = e_raw.astype(np.float64)
                    out["pos_error"] = pe
                    out["reward_hover"] = np.exp(-cfg.reward_k * pe * pe)
                    out["reward_action"] = -0.001 * asq.astype(np.float64)
                    out["reward"] = out["reward_hover"]            # Q2: action term dropped
                    out_xy = (np.abs(pos[:, 0]) > cfg.pos_limit_xy) | (np.abs(pos[:, 1]) > cfg.pos_limit_xy)
                    out_z = (pos[:, 2] < F32(cfg.z_low)) | (pos[:, 2] > F32(cfg.z_high))
                    out["done"] = (out_xy | out_z).astype(F32)
                else:
                    pe = e_raw.astype(np.float64)
                    out["pos_error"] = pe
                    out["reward_hover"] = np.exp(-pe * pe); out["reward_action"] = np.zeros(n)
                    out["reward"] = out["reward_hover"]
                    out["done"] = np.zeros(n, F32)                 # jax_mjx_quad_env.py:170-172
                    out["truncated"] = (s["step_count"] >= cfg.max_episode_steps).astype(F32)
                out["obs"] = obs.astype(F32)
        return out

    # ------------------------------------------------------------------ reset
    def reset(self, s, mask=None):
        """Philox reset of the masked envs, in place; mirrors qs_reset / reset_env."""
        cfg = self.cfg
        n = s["qpos"].shape[0]
        m = np.ones(n, bool) if mask is None else np.asarray(mask, bool)
        ids = (np.arange(n, dtype=np.uint32) + np.uint32(cfg.env_id_offset))[m]
        k = int(m.sum())
        if k == 0:
            return
        s["step_count"][m] = 0; s["ep_steps"][m] = 0; s["done_prev"][m] = 0; s["voltage"][m] = cfg.v_nominal
        s["rate_int"][m] = 0.0; s["prev_action"][m] = 0.0          # rate_wrapper.py:108-111, hover_env.py:211
        qpos = np.zeros((k, 11)); qvel = np.zeros((k, 10)); qpos[:, 3] = 1.0
        if _gym(cfg.mode):
            if cfg.waypoint_mode:
                shape = ids % np.uint32(len(cfg.waypoints))
                for j in range(k):
                    wp = cfg.waypoints[int(shape[j])]
                    qpos[j, 0:3] = wp[0].astype(F32)
                    idx = 1 % len(wp)
                    s["wp_idx"][np.flatnonzero(m)[j]] = idx
                    s["target"][np.flatnonzero(m)[j]] = wp[idx].astype(F32)
            else:
                nb = 4 if cfg.mode == Q.MODE_HOVER_GYM else 3
                raw = philox.draw_blocks(cfg.seed, ids, s["episode"][m], nb, philox.STREAM_RESET)
                s12 = np.stack([philox.uniform(raw[:, i], cfg.init_lo[i], cfg.init_hi[i]) for i in range(12)], axis=1)
                qpos[:, 0:3] = s12[:, 0:3]
                quat = Rotation.from_euler("xyz", s12[:, 3:6].astype(np.float64)).as_quat()   # utils/state.py:59
                qpos[:, 3:7] = quat[:, [3, 0, 1, 2]]
                qvel[:, 0:6] = s12[:, 6:12]
                if cfg.mode == Q.MODE_HOVER_GYM:
                    s["target"][m] = np.stack([philox.uniform(raw[:, 12 + i], cfg.target_lo[i], cfg.target_hi[i])
                                               for i in range(3)], axis=1)
                else:
                    s["target"][m] = s12[:, 0:3]               # Q7
        elif cfg.mode == Q.MODE_MJX_PLAYGROUND:
            qpos[:, 0:3] = 0.0
        else:
            raw = philox.draw_blocks(cfg.seed, ids, s["episode"][m], 6, philox.STREAM_RESET)
            noise = np.stack([philox.uniform(raw[:, i], -cfg.reset_noise, cfg.reset_noise) for i in range(21)], axis=1)
            q0 = np.zeros(11, F32); q0[2] = cfg.reset_z; q0[3] = 1.0
            q = (q0[None, :] + noise[:, :11]).astype(F32)
            if cfg.mode == Q.MODE_MJX_BRAX:
                nrm = np.sqrt(np.sum(q[:, 3:7] * q[:, 3:7], axis=1, dtype=F32))
                q[:, 3:7] = q[:, 3:7] / (nrm + F32(1e-8))[:, None]
            qpos[:] = q
            qvel[:] = noise[:, 11:]
        s["qpos"][m] = qpos; s["qvel"][m] = qvel

    # ------------------------------------------------------------------ waypoint advance
    def waypoint_advance(self, s):
        """evaluate.py:540-557.  Returns lap-completed mask."""
        cfg = self.cfg
        n = s["qpos"].shape[0]
        ids = np.arange(n, dtype=np.uint32) + np.uint32(cfg.env_id_offset)
        lap = np.zeros(n, bool)
        pos32 = s["qpos"][:, 0:3].astype(F32)
        for i in range(n):
            wp = cfg.waypoints[int(ids[i] % len(cfg.waypoints))]
            d = pos32[i].astype(np.float64) - wp[s["wp_idx"][i]]
            dist = np.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
            if dist < cfg.wp_reach_radius:
                s["wp_reached"][i] += 1
                s["wp_idx"][i] = (s["wp_idx"][i] + 1) % len(wp)
                if s["wp_idx"][i] == 0:
                    s["laps"][i] += 1
                    lap[i] = True
                else:
                    s["target"][i] = wp[s["wp_idx"][i]].astype(F32)
        return lap

    # ------------------------------------------------------------------ full step
    def step(self, s, action, first=None):
        """One engine step (env.step + configured wrappers), in place on `s`.  Returns the output dict."""
        cfg = self.cfg
        n = s["qpos"].shape[0]
        if _brax(cfg.mode) and cfg.auto_reset == Q.RESET_RESTORE_FIRST:
            s["ep_steps"][s["done_prev"] != 0] = 0
        env_action = action
        if _gym(cfg.mode) and cfg.rate_wrapper:
            env_action = self.rate_to_torque(s, action)
            s["prev_action"] = np.asarray(action, dtype=F32).copy()       # rate_wrapper.py:100-104
        ctrl, s["voltage"] = self.action_to_ctrl(env_action, s["voltage"])
        with np.errstate(all="ignore"):
            s["qpos"], s["qvel"] = self.pipe.step(s["qpos"], s["qvel"], ctrl)
        s["step_count"] = s["step_count"] + 1
        out = self.evaluate(s, action)
        out["finished"] = np.zeros(n, bool)
        out["terminal_obs"] = np.full_like(out["obs"], np.nan)
        if _gym(cfg.mode):
            lap = self.waypoint_advance(s) if cfg.waypoint_mode else np.zeros(n, bool)
            fin = (out["done"] != 0) | (out["truncated"] != 0) | lap
            out["finished"] = fin
            out["terminal_obs"][fin] = out["obs"][fin]
            if cfg.auto_reset == Q.RESET_RESAMPLE and fin.any():
                s["episode"][fin] += 1
                self.reset(s, fin)
                ev = self.evaluate(s, None)
                out["obs"][fin] = ev["obs"][fin]
        elif _brax(cfg.mode):
            if cfg.episode_length > 0:
                s["ep_steps"] = s["ep_steps"] + 1
                over = s["ep_steps"] >= cfg.episode_length
                out["truncated"] = np.where(over, 1.0 - out["done"], 0.0).astype(F32)
                out["done"] = np.where(over, 1.0, out["done"]).astype(F32)
            fin = out["done"] != 0
            out["finished"] = fin
            out["terminal_obs"][fin] = out["obs"][fin]
            if cfg.auto_reset == Q.RESET_RESTORE_FIRST:
                if fin.any():
                    s["qpos"][fin] = first["qpos"][fin]; s["qvel"][fin] = first["qvel"][fin]
                    ev = self.evaluate(s, None)
                    out["obs"][fin] = ev["obs"][fin]
                s["done_prev"] = out["done"].copy()
        else:
            out["finished"] = out["truncated"] != 0
            out["terminal_obs"][out["finished"]] = out["obs"][out["finished"]]
        return out
