"""Env configuration -> the C ``QsParams`` block (include/quadsim_abi.h).

``EnvConfig`` gathers every constant of the reference's env variants with the file:line
it comes from; ``pack_params`` merges it with the model constants (``model.QuadConstants``)
into the ctypes mirror of ``QsParams`` that is handed to ``qs_create`` and from there,
by value, to every kernel.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field, replace

import numpy as np

from .model import QuadConstants

MODE_MJX_BRAX, MODE_HOVER_GYM, MODE_TRAJ_GYM, MODE_HOVER_BRAX, MODE_MJX_PLAYGROUND = range(5)
MODE_NAMES = {"mjx_brax": 0, "hover_gym": 1, "traj_gym": 2, "hover_brax": 3, "mjx_playground": 4}
RESET_NONE, RESET_RESTORE_FIRST, RESET_RESAMPLE = 0, 1, 2
NPLANES = 40
MAX_WP = 64
MAX_SHAPES = 4

# utils/drone_config.py:9-22
MAX_MOTOR_THRUST = 13.0
ARM_LENGTH = 0.039799
YAW_TORQUE_COEFF = 0.0201
MAX_TORQUE = 0.5
MAX_TOTAL_THRUST = 4 * MAX_MOTOR_THRUST


class QsParams(C.Structure):
    _fields_ = [
        ("dt", C.c_float), ("gz", C.c_float), ("mass", C.c_float), ("inv_mass", C.c_float),
        ("com", C.c_float * 3), ("I_C", C.c_float * 9), ("Ieff_inv", C.c_float * 9),
        ("rotor_J", C.c_float * 4), ("rotor_invJs", C.c_float * 4), ("rotor_rho", C.c_float * 4),
        ("rotor_damp", C.c_float * 4), ("rotor_r", C.c_float * 12), ("rotor_d", C.c_float * 12),
        ("wrench", C.c_float * 24), ("ctrl_lo", C.c_float * 4), ("ctrl_hi", C.c_float * 4),
        ("base_lin_visc", C.c_float), ("base_ang_visc", C.c_float),
        ("base_lin_quad", C.c_float * 3), ("base_ang_quad", C.c_float * 3),
        ("rot_lin_visc", C.c_float * 4), ("rot_lin_quad_ax", C.c_float * 4), ("rot_lin_quad_lat", C.c_float * 4),
        ("rot_ang_visc", C.c_float * 4), ("rot_ang_quad_ax", C.c_float * 4), ("rot_ang_quad_lat", C.c_float * 4),
        ("act_lo", C.c_float * 4), ("act_hi", C.c_float * 4), ("mix_inv", C.c_float * 16),
        ("max_motor_thrust", C.c_float), ("pre_clip_action", C.c_int32),
        ("battery", C.c_int32), ("v_nominal", C.c_float), ("v_min", C.c_float),
        ("v_drop_base", C.c_float), ("v_drop_load", C.c_float),
        ("mode", C.c_int32), ("obs_dim", C.c_int32),
        ("obs_lo", C.c_float * 12), ("obs_scale", C.c_float * 12),
        ("term_lo", C.c_float * 12), ("term_hi", C.c_float * 12),
        ("pos_limit_xy", C.c_float), ("z_low", C.c_float), ("z_high", C.c_float), ("vel_limit", C.c_float),
        ("reward_k", C.c_float), ("action_penalty", C.c_float), ("fixed_target", C.c_float * 3),
        ("max_episode_steps", C.c_int32), ("episode_length", C.c_int32), ("auto_reset", C.c_int32),
        ("init_lo", C.c_float * 12), ("init_hi", C.c_float * 12),
        ("target_lo", C.c_float * 3), ("target_hi", C.c_float * 3),
        ("reset_noise", C.c_float), ("reset_z", C.c_float),
        ("seed_lo", C.c_uint32), ("seed_hi", C.c_uint32), ("env_id_offset", C.c_uint32), ("philox_key", C.c_uint32 * 20),
        ("waypoint_mode", C.c_int32), ("wp_num_shapes", C.c_int32), ("wp_count", C.c_int32 * MAX_SHAPES),
        ("wp_reach_radius", C.c_float),
        ("rate_wrapper", C.c_int32), ("rate_max", C.c_float), ("rate_kd", C.c_float * 3),
        ("rate_inertia", C.c_float * 3), ("rate_ki", C.c_float), ("rate_imax", C.c_float), ("max_torque", C.c_float),
        ("spline_duration", C.c_float), ("traj_center_lo", C.c_float * 3), ("traj_center_hi", C.c_float * 3),
        ("traj_amp", C.c_float * 3),
        ("reserved", C.c_int32 * 5),
    ]


class QsPolicyDesc(C.Structure):
    _fields_ = [("obs_dim", C.c_int32), ("hidden", C.c_int32), ("act_dim", C.c_int32), ("dist", C.c_int32),
                ("deterministic", C.c_int32), ("bootstrap_gamma", C.c_float), ("tensor_cores", C.c_int32),
                ("sample_seed", C.c_int32)]


_PI32 = float(np.float32(np.pi))


def _f32(x):
    return np.asarray(x, dtype=np.float32)


@dataclass(frozen=True)
class EnvConfig:
    """All env-semantics constants; defaults = HoverEnv (envs/hover_env.py)."""
    mode: int = MODE_HOVER_GYM
    # action path (hover_env.py:60-65; train_brax_ppo.py:229-230)
    act_lo: tuple = (0.0, -MAX_TORQUE, -MAX_TORQUE, -MAX_TORQUE)
    act_hi: tuple = (MAX_TOTAL_THRUST, MAX_TORQUE, MAX_TORQUE, MAX_TORQUE)
    arm_length: float = ARM_LENGTH
    yaw_coeff: float = YAW_TORQUE_COEFF
    max_motor_thrust: float = MAX_MOTOR_THRUST
    pre_clip_action: bool = False                 # Q1 (train_brax_ppo.py:134,310)
    # battery sag (hover_env.py:18-20,102-109); the north-star default leaves it off
    battery: bool = False
    v_nominal: float = 8.4
    v_min: float = 7.6
    v_drop_base: float = 0.01
    v_drop_load: float = 0.08
    # observation normalisation bounds (hover_env.py:36-39), float32 like the reference's Box
    obs_lo: tuple = (-4, -4, -2, -_PI32, -_PI32, -_PI32, -10, -10, -10,
                     -float(np.float32(6 * np.pi)), -float(np.float32(6 * np.pi)), -float(np.float32(6 * np.pi)))
    obs_hi: tuple = (4, 4, 2, _PI32, _PI32, _PI32, 10, 10, 10,
                     float(np.float32(6 * np.pi)), float(np.float32(6 * np.pi)), float(np.float32(6 * np.pi)))
    # termination box (hover_env.py:54-57), inclusive
    term_lo: tuple = (-2, -2, 0.0, -_PI32, -_PI32, -_PI32, -10, -10, -10,
                      -float(np.float32(6 * np.pi)), -float(np.float32(6 * np.pi)), -float(np.float32(6 * np.pi)))
    term_hi: tuple = (2, 2, 2, _PI32, _PI32, _PI32, 10, 10, 10,
                      float(np.float32(6 * np.pi)), float(np.float32(6 * np.pi)), float(np.float32(6 * np.pi)))
    # brax-mode limits (train_brax_ppo.py:184-191)
    pos_limit_xy: float = 3.0
    z_low: float = 0.02
    z_high: float = 4.0
    vel_limit: float = 20.0
    reward_k: float = 1.0
    action_penalty: float = 0.0
    fixed_target: tuple = (0.0, 0.0, 1.0)
    traj_duration_seconds: float = 5.0
    # episodes
    max_episode_steps: int = 512                  # hover_env.py:16
    episode_length: int = 0                       # Brax EpisodeWrapper (train_brax_ppo.py:436); 0 = off
    auto_reset: int = RESET_NONE
    # reset distribution (hover_env.py:42-51)
    init_lo: tuple = (-1.5, -1.5, 0.1, -0.3, -0.3, -0.3, -0.5, -0.5, -0.5, -0.5, -0.5, -0.5)
    init_hi: tuple = (1.5, 1.5, 1.5, 0.3, 0.3, 0.3, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5)
    target_lo: tuple = (-1.5, -1.5, 0.3)
    target_hi: tuple = (1.5, 1.5, 1.8)
    reset_noise: float = 0.01                     # train_brax_ppo.py:260-261
    reset_z: float = 1.0                          # train_brax_ppo.py:251
    seed: int = 0
    env_id_offset: int = 0
    # waypoint tracking (evaluate.py:440-442)
    waypoint_mode: bool = False
    waypoints: tuple = ()                         # tuple of (n_i, 3) float64 arrays, one per shape
    wp_reach_radius: float = 0.25
    # RateControlWrapper (envs/rate_wrapper.py:16-23,58; pid_gains.json "rate_wrapper"; train.py:31 default wrapper)
    rate_wrapper: bool = False
    rate_max_deg: float = 360.0
    rate_kd: tuple = (26.0, 26.0, 18.0)
    rate_ki: float = 0.025
    rate_imax: float = 0.01
    rate_inertia: tuple = (4.16e-4, 4.23e-4, 5.37e-4)   # IXX, IYY, IZZ (utils/drone_config.py:15-17)
    # TrajectoryFollowEnv spline reference (envs/trajectory_follow_env.py:25,55-58,203)
    spline_duration: float | None = 30.0          # traj_duration_seconds; None: sample times = arange(N) * dt
    dt_nominal: float = 0.01                      # model timestep, used only when spline_duration is None
    traj_center_lo: tuple = (-1.0, -1.0, 0.4)
    traj_center_hi: tuple = (1.0, 1.0, 1.4)
    traj_amp: tuple = (0.6, 0.6, 0.4)

    # ---- the reference's five env variants -------------------------------------------
    @staticmethod
    def hover_gym(**kw):
        """envs/hover_env.py HoverEnv; battery sag on = exact reference semantics."""
        return replace(EnvConfig(mode=MODE_HOVER_GYM, battery=True), **kw)

    @staticmethod
    def north_star(**kw):
        """HoverEnv semantics without battery sag, VecEnv-style Philox auto-reset (BASELINE north_star)."""
        return replace(EnvConfig(mode=MODE_HOVER_GYM, battery=False, auto_reset=RESET_RESAMPLE), **kw)

    @staticmethod
    def traj_gym(**kw):
        """envs/trajectory_follow_env.py TrajectoryFollowEnv (:24-28,60-63)."""
        tl = (-3, -3, 0.0) + EnvConfig().term_lo[3:]
        th = (3, 3, 3) + EnvConfig().term_hi[3:]
        return replace(EnvConfig(mode=MODE_TRAJ_GYM, battery=True, v_nominal=16.8, v_min=13.2,
                                 max_episode_steps=2048, term_lo=tl, term_hi=th), **kw)

    @staticmethod
    def mjx_brax(**kw):
        """train_brax_ppo.py:179-368 JaxMJXQuadBraxEnv."""
        return replace(EnvConfig(mode=MODE_MJX_BRAX, pre_clip_action=True, action_penalty=0.001,
                                 max_episode_steps=500, reset_z=1.0), **kw)

    @staticmethod
    def hover_brax(**kw):
        """train_brax_ppo.py:39-176 QuadHoverBraxEnv (init_q has z = 0)."""
        return replace(EnvConfig(mode=MODE_HOVER_BRAX, pre_clip_action=True, reward_k=2.0, reset_z=0.0), **kw)

    @staticmethod
    def mjx_playground(**kw):
        """envs/jax_mjx_quad_env.py JaxMJXQuadEnv."""
        return replace(EnvConfig(mode=MODE_MJX_PLAYGROUND, max_episode_steps=500), **kw)

    @staticmethod
    def waypoint_eval(waypoints, **kw):
        """evaluate.py:440-557 evaluate_trajectory: HoverEnv(max_episode_steps=5000) + waypoint advance."""
        return replace(EnvConfig(mode=MODE_HOVER_GYM, battery=True, max_episode_steps=5000,
                                 waypoint_mode=True, waypoints=tuple(np.asarray(w, dtype=np.float64) for w in waypoints)),
                       **kw)

    @property
    def obs_dim(self):
        return 12 if self.mode in (MODE_HOVER_GYM, MODE_TRAJ_GYM) else 21

    def position_bounds(self):
        """(lo, hi) of the base position while an episode is alive -- what the loader's contact guard
        (model.check_contacts) needs: the gym modes' termination box (hover_env.py:54-57), the Brax modes' limits
        (train_brax_ppo.py:184-191), (None, None) for JaxMJXQuadEnv, which never terminates."""
        if self.mode in (MODE_HOVER_GYM, MODE_TRAJ_GYM):
            return tuple(self.term_lo[:3]), tuple(self.term_hi[:3])
        if self.mode in (MODE_MJX_BRAX, MODE_HOVER_BRAX):
            return (-self.pos_limit_xy, -self.pos_limit_xy, self.z_low), (self.pos_limit_xy, self.pos_limit_xy, self.z_high)
        return None, None

    def position_overshoot(self, dt: float) -> float:
        """How far the base can travel in the one step that ends an episode: (v_max + a_max dt) dt with v_max the
        velocity bound of the mode (hover_brax has none: 20 m/s is assumed) and a_max = full thrust / mass + g."""
        v = max(abs(float(self.term_lo[6])), abs(float(self.term_hi[6]))) if self.mode in (MODE_HOVER_GYM, MODE_TRAJ_GYM) else self.vel_limit
        a = 4.0 * self.max_motor_thrust / 0.2227 + 9.81
        return (v + a * dt) * dt

    def mixer(self):
        """A and A^-1 (hover_env.py:94-100), float64."""
        l, k = self.arm_length, self.yaw_coeff
        A = np.array([[1, 1, 1, 1], [-l, -l, l, l], [-l, l, l, -l], [k, -k, k, -k]], dtype=np.float64)
        return A, np.linalg.inv(A)

    def target_table(self):
        """train_brax_ppo.py:358-364 in float32 (jp default dtype): [max_episode_steps][3]."""
        N = int(self.max_episode_steps)
        t = np.linspace(np.float32(0.0), np.float32(self.traj_duration_seconds), N, dtype=np.float32)
        center = _f32([0.0, 0.0, 1.0]); amp = _f32([0.5, 0.5, 0.2]); freq = _f32([0.2, 0.15, 0.1])
        ang = (np.float32(2.0) * np.float32(np.pi) * freq)[None, :] * t[:, None]
        return (center + amp * np.sin(ang.astype(np.float32))).astype(np.float32)

    def waypoint_table(self):
        tab = np.zeros((max(len(self.waypoints), 1), MAX_WP, 3), dtype=np.float64)
        for s, w in enumerate(self.waypoints):
            if len(w) > MAX_WP:
                raise ValueError("too many waypoints")
            tab[s, :len(w)] = w
        return tab


def pack_params(c: QuadConstants, cfg: EnvConfig) -> QsParams:
    P = QsParams()

    def put(name, arr):
        a = np.asarray(arr, dtype=np.float64).ravel()
        fld = getattr(P, name)
        if len(fld) != a.size:
            raise ValueError(f"{name}: expected {len(fld)} values, got {a.size}")
        for i, v in enumerate(a):
            fld[i] = float(v)

    P.dt, P.gz, P.mass, P.inv_mass = c.dt, c.gz, c.mass, 1.0 / c.mass
    put("com", c.com); put("I_C", c.I_C); put("Ieff_inv", c.Ieff_inv)
    put("rotor_J", c.rotor_J); put("rotor_invJs", 1.0 / c.rotor_Js); put("rotor_rho", c.rotor_J / c.rotor_Js)
    put("rotor_damp", c.rotor_damping); put("rotor_r", c.rotor_r); put("rotor_d", c.rotor_r - c.com[None, :])
    put("wrench", c.wrench)
    big = 3.0e38
    put("ctrl_lo", np.clip(c.ctrl_lo, -big, big)); put("ctrl_hi", np.clip(c.ctrl_hi, -big, big))
    P.base_lin_visc, P.base_ang_visc = c.base_lin_visc, c.base_ang_visc
    put("base_lin_quad", c.base_lin_quad); put("base_ang_quad", c.base_ang_quad)
    put("rot_lin_visc", c.rot_lin_visc); put("rot_lin_quad_ax", c.rot_lin_quad_ax)
    put("rot_lin_quad_lat", c.rot_lin_quad_lat); put("rot_ang_visc", c.rot_ang_visc)
    put("rot_ang_quad_ax", c.rot_ang_quad_ax); put("rot_ang_quad_lat", c.rot_ang_quad_lat)
    put("act_lo", cfg.act_lo); put("act_hi", cfg.act_hi)
    put("mix_inv", cfg.mixer()[1])
    P.max_motor_thrust = cfg.max_motor_thrust
    P.pre_clip_action = int(cfg.pre_clip_action)
    P.battery = int(cfg.battery)
    P.v_nominal, P.v_min, P.v_drop_base, P.v_drop_load = cfg.v_nominal, cfg.v_min, cfg.v_drop_base, cfg.v_drop_load
    P.mode, P.obs_dim = cfg.mode, cfg.obs_dim
    lo = _f32(cfg.obs_lo); hi = _f32(cfg.obs_hi)
    put("obs_lo", lo); put("obs_scale", 2.0 / (hi.astype(np.float64) - lo.astype(np.float64)))
    put("term_lo", _f32(cfg.term_lo)); put("term_hi", _f32(cfg.term_hi))
    P.pos_limit_xy, P.z_low, P.z_high, P.vel_limit = cfg.pos_limit_xy, cfg.z_low, cfg.z_high, cfg.vel_limit
    P.reward_k, P.action_penalty = cfg.reward_k, cfg.action_penalty
    put("fixed_target", cfg.fixed_target)
    P.max_episode_steps, P.episode_length, P.auto_reset = cfg.max_episode_steps, cfg.episode_length, cfg.auto_reset
    put("init_lo", cfg.init_lo); put("init_hi", cfg.init_hi)
    put("target_lo", cfg.target_lo); put("target_hi", cfg.target_hi)
    P.reset_noise, P.reset_z = cfg.reset_noise, cfg.reset_z
    P.seed_lo, P.seed_hi = cfg.seed & 0xFFFFFFFF, (cfg.seed >> 32) & 0xFFFFFFFF
    P.env_id_offset = cfg.env_id_offset
    for r in range(10):
        P.philox_key[2 * r] = (P.seed_lo + r * 0x9E3779B9) & 0xFFFFFFFF
        P.philox_key[2 * r + 1] = (P.seed_hi + r * 0xBB67AE85) & 0xFFFFFFFF
    P.waypoint_mode = int(cfg.waypoint_mode)
    P.wp_num_shapes = len(cfg.waypoints)
    for s, w in enumerate(cfg.waypoints):
        P.wp_count[s] = len(w)
    P.wp_reach_radius = cfg.wp_reach_radius
    P.rate_wrapper = int(cfg.rate_wrapper)
    P.rate_max = math.radians(cfg.rate_max_deg)
    put("rate_kd", cfg.rate_kd); put("rate_inertia", cfg.rate_inertia)
    P.rate_ki, P.rate_imax, P.max_torque = cfg.rate_ki, cfg.rate_imax, MAX_TORQUE
    P.spline_duration = float(cfg.spline_duration) if cfg.spline_duration is not None else 0.0
    put("traj_center_lo", cfg.traj_center_lo); put("traj_center_hi", cfg.traj_center_hi); put("traj_amp", cfg.traj_amp)
    return P


def params_to_dict(P: QsParams) -> dict:
    out = {}
    for name, ty in P._fields_:
        v = getattr(P, name)
        out[name] = list(v) if hasattr(v, "__len__") else v
    return out
