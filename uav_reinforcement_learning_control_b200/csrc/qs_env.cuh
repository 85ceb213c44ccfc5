// qs_env.cuh -- env semantics around the physics step, one env per thread, state in registers.
//
// One template parameter (MODE) selects which reference env is reproduced:
//   QS_MODE_MJX_BRAX       JaxMJXQuadBraxEnv.step   train_brax_ppo.py:307-356
//   QS_MODE_HOVER_GYM      HoverEnv.step            envs/hover_env.py:159-198
//   QS_MODE_TRAJ_GYM       TrajectoryFollowEnv.step envs/trajectory_follow_env.py:146-174
//   QS_MODE_HOVER_BRAX     QuadHoverBraxEnv.step    train_brax_ppo.py:131-173
//   QS_MODE_MJX_PLAYGROUND JaxMJXQuadEnv.step       envs/jax_mjx_quad_env.py:124-158
// and the vectorising wrappers' episode logic (Brax EpisodeWrapper/AutoResetWrapper, SB3
// VecEnv auto-reset) is fused in as configured by QsParams.episode_length / auto_reset.
#pragma once

#include "qs_dynamics.cuh"
#include "qs_philox.cuh"

namespace qs {

template <int MODE> struct ModeTraits {
    static constexpr bool kGym = (MODE == QS_MODE_HOVER_GYM || MODE == QS_MODE_TRAJ_GYM);
    static constexpr bool kBrax = (MODE == QS_MODE_MJX_BRAX || MODE == QS_MODE_HOVER_BRAX);
    static constexpr bool kTable = (MODE == QS_MODE_MJX_BRAX || MODE == QS_MODE_MJX_PLAYGROUND);
    static constexpr int kObsDim = kGym ? 12 : 21;
};

// Compile-time feature gates.  FeatAll (default) leaves every optional stage under its run-time QsParams flag;
// FeatLean folds them away for the plain north-star configuration (no battery sag, no rate wrapper, no waypoint
// table, no action pre-clip), which the step launcher selects when the handle's flags allow it.
struct FeatAll { static constexpr bool kBattery = true, kRate = true, kWaypoint = true, kPreClip = true; };
struct FeatLean { static constexpr bool kBattery = false, kRate = false, kWaypoint = false, kPreClip = false; };

// Per-env registers.  Which members are live (and which planes are touched in HBM) depends on MODE.
struct Env {
    Body b;
    float target[3];
    int32_t step_count;
    float voltage;
    uint32_t episode;
    int32_t ep_steps;
    int32_t wp_idx, wp_reached, laps;
    float done_prev;
    float rate_int[3];     // RateControlWrapper integral state
    float prev_action[4];  // last policy action (what RelPosActWrapper appends to the observation)
};

struct StepOut {
    float reward, done, truncated;
    float pos_error, reward_hover, reward_action;
    bool finished;        // episode ended this step (terminal_obs is meaningful)
    bool needs_reset;     // DEFER_RESET only: caller must run reset_env + compute_obs for this env
};

// device-resident tables owned by the engine handle
struct Tables {
    const float* __restrict__ target;    // [max_episode_steps][3]   (mjx modes)
    const double* __restrict__ waypoints;  // [shapes][QS_MAX_WP][3] (waypoint mode)
};

// bit casts for the int32/uint32 planes of the float32 state array
QS_HD int32_t f2i_(float f) { union { float f; int32_t i; } v; v.f = f; return v.i; }
QS_HD uint32_t f2u_(float f) { union { float f; uint32_t u; } v; v.f = f; return v.u; }
QS_HD float i2f_(int32_t i) { union { float f; int32_t i; } v; v.i = i; return v.f; }
QS_HD float u2f_(uint32_t u) { union { float f; uint32_t u; } v; v.u = u; return v.f; }

// ---------------------------------------------------------------------------------------
// planar state <-> registers.  Only the planes the MODE / flags need are touched, which is
// what keeps a hover step at 27 words in + 27 words out per env (include/quadsim_abi.h).
// ---------------------------------------------------------------------------------------
template <int MODE, class F = FeatAll>
QS_HD void load_env(const QsParams& P, const float* __restrict__ st, int n, int i, Env& e, bool for_reset = false) {
    using M = ModeTraits<MODE>;
    const float* s = st + i;
    e.b.p[0] = s[0 * (size_t)n]; e.b.p[1] = s[1 * (size_t)n]; e.b.p[2] = s[2 * (size_t)n];
    e.b.q[0] = s[3 * (size_t)n]; e.b.q[1] = s[4 * (size_t)n]; e.b.q[2] = s[5 * (size_t)n]; e.b.q[3] = s[6 * (size_t)n];
#pragma unroll
    for (int k = 0; k < 4; ++k) e.b.th[k] = s[(7 + k) * (size_t)n];
    e.b.v[0] = s[11 * (size_t)n]; e.b.v[1] = s[12 * (size_t)n]; e.b.v[2] = s[13 * (size_t)n];
    e.b.w[0] = s[14 * (size_t)n]; e.b.w[1] = s[15 * (size_t)n]; e.b.w[2] = s[16 * (size_t)n];
#pragma unroll
    for (int k = 0; k < 4; ++k) e.b.s[k] = s[(17 + k) * (size_t)n];
    e.target[0] = e.target[1] = e.target[2] = 0.f;
    e.step_count = 0; e.voltage = P.v_nominal; e.ep_steps = 0;
    e.episode = for_reset ? f2u_(s[26 * (size_t)n]) : 0u;   // Philox counter word; brax modes only need it in reset
    e.wp_idx = 0; e.wp_reached = 0; e.laps = 0; e.done_prev = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) e.rate_int[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) e.prev_action[k] = 0.f;
    if constexpr (M::kGym) {
        e.target[0] = s[21 * (size_t)n]; e.target[1] = s[22 * (size_t)n]; e.target[2] = s[23 * (size_t)n];
        e.step_count = f2i_(s[24 * (size_t)n]);
        if (F::kBattery && P.battery) e.voltage = s[25 * (size_t)n];
        e.episode = f2u_(s[26 * (size_t)n]);
        if (F::kRate && P.rate_wrapper) {
#pragma unroll
            for (int k = 0; k < 3; ++k) e.rate_int[k] = s[(32 + k) * (size_t)n];
        }
        if (F::kWaypoint && P.waypoint_mode) {
            e.wp_idx = f2i_(s[28 * (size_t)n]);
            e.wp_reached = f2i_(s[29 * (size_t)n]);
            e.laps = f2i_(s[30 * (size_t)n]);
        }
    } else {
        if constexpr (MODE != QS_MODE_HOVER_BRAX) e.step_count = f2i_(s[24 * (size_t)n]);
        if constexpr (M::kBrax) {
            if (P.episode_length > 0 || P.auto_reset != QS_RESET_NONE) {
                e.ep_steps = f2i_(s[27 * (size_t)n]);
                e.done_prev = s[31 * (size_t)n];
            }
        }
    }
}

template <int MODE, class F = FeatAll>
QS_HD void store_env(const QsParams& P, float* __restrict__ st, int n, int i, const Env& e) {
    using M = ModeTraits<MODE>;
    float* s = st + i;
    s[0 * (size_t)n] = e.b.p[0]; s[1 * (size_t)n] = e.b.p[1]; s[2 * (size_t)n] = e.b.p[2];
    s[3 * (size_t)n] = e.b.q[0]; s[4 * (size_t)n] = e.b.q[1]; s[5 * (size_t)n] = e.b.q[2]; s[6 * (size_t)n] = e.b.q[3];
#pragma unroll
    for (int k = 0; k < 4; ++k) s[(7 + k) * (size_t)n] = e.b.th[k];
    s[11 * (size_t)n] = e.b.v[0]; s[12 * (size_t)n] = e.b.v[1]; s[13 * (size_t)n] = e.b.v[2];
    s[14 * (size_t)n] = e.b.w[0]; s[15 * (size_t)n] = e.b.w[1]; s[16 * (size_t)n] = e.b.w[2];
#pragma unroll
    for (int k = 0; k < 4; ++k) s[(17 + k) * (size_t)n] = e.b.s[k];
    if constexpr (M::kGym) {
        s[21 * (size_t)n] = e.target[0]; s[22 * (size_t)n] = e.target[1]; s[23 * (size_t)n] = e.target[2];
        s[24 * (size_t)n] = i2f_(e.step_count);
        if (F::kBattery && P.battery) s[25 * (size_t)n] = e.voltage;
        s[26 * (size_t)n] = u2f_(e.episode);
        if (F::kRate && P.rate_wrapper) {
#pragma unroll
            for (int k = 0; k < 3; ++k) s[(32 + k) * (size_t)n] = e.rate_int[k];
#pragma unroll
            for (int k = 0; k < 4; ++k) s[(35 + k) * (size_t)n] = e.prev_action[k];
        }
        if (F::kWaypoint && P.waypoint_mode) {
            s[28 * (size_t)n] = i2f_(e.wp_idx);
            s[29 * (size_t)n] = i2f_(e.wp_reached);
            s[30 * (size_t)n] = i2f_(e.laps);
        }
    } else {
        if constexpr (MODE != QS_MODE_HOVER_BRAX) s[24 * (size_t)n] = i2f_(e.step_count);
        if constexpr (M::kBrax) {
            if (P.episode_length > 0 || P.auto_reset != QS_RESET_NONE) {
                s[27 * (size_t)n] = i2f_(e.ep_steps);
                s[31 * (size_t)n] = e.done_prev;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// action -> motor forces   (hover_env.py:169-177; train_brax_ppo.py:309-314)
// ---------------------------------------------------------------------------------------
template <class F = FeatAll>
QS_HD void action_to_ctrl(const QsParams& P, const float a[4], float& voltage, float ctrl[4]) {
    float u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        // (a + 1) / 2 * (hi - lo) + lo
        u[i] = fma_((a[i] + 1.0f) * 0.5f, P.act_hi[i] - P.act_lo[i], P.act_lo[i]);
        if (F::kPreClip && P.pre_clip_action) u[i] = clamp_(u[i], P.act_lo[i], P.act_hi[i]);   // Q1
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float Fk = fma_(P.mix_inv[4 * k], u[0], fma_(P.mix_inv[4 * k + 1], u[1],
                         fma_(P.mix_inv[4 * k + 2], u[2], P.mix_inv[4 * k + 3] * u[3])));
        ctrl[k] = clamp_(Fk, 0.0f, P.max_motor_thrust);
    }
    if (F::kBattery && P.battery) {
        // hover_env.py:102-109,174-176
        const float scale = clamp_(voltage / P.v_nominal, 0.0f, 1.0f);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ctrl[k] = clamp_(ctrl[k] * scale, 0.0f, P.max_motor_thrust * scale);
            sum += ctrl[k];
        }
        const float load = (0.25f * sum) / fmaxf(P.max_motor_thrust, 1e-6f);
        const float dV = fma_(P.v_drop_load, load, P.v_drop_base) * P.dt;
        voltage = clamp_(voltage - dV, P.v_min, P.v_nominal);
    }
}

// ---------------------------------------------------------------------------------------
// RateControlWrapper.action (envs/rate_wrapper.py:69-98): [thrust, roll/pitch/yaw rate] in [-1,1] ->
// [thrust, tau_x, tau_y, tau_z] in [-1,1] through a per-axis PI loop on the body rates.
// ---------------------------------------------------------------------------------------
QS_HD void rate_to_torque(const QsParams& P, const float a[4], const float w[3], float rate_int[3], float out[4]) {
    out[0] = a[0];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float err = fma_(a[1 + k], P.rate_max, -w[k]);
        const float tau_p = P.rate_inertia[k] * P.rate_kd[k] * err;
        rate_int[k] = clamp_(fma_(P.rate_ki * P.dt, err, rate_int[k]), -P.rate_imax, P.rate_imax);
        out[1 + k] = clamp_((tau_p + rate_int[k]) / P.max_torque, -1.0f, 1.0f);
    }
}

// ---------------------------------------------------------------------------------------
// Euler angles, scipy Rotation.as_euler('xyz') convention (utils/state.py:42):
// R = Rz(yaw) Ry(pitch) Rx(roll)
// ---------------------------------------------------------------------------------------
QS_HD void quat_to_rpy(const float q[4], float rpy[3]) {
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    rpy[0] = atan2_(2.f * fma_(y, z, w * x), fma_(-2.f, fma_(x, x, y * y), 1.f));
    rpy[1] = asin_unit_(clamp_(2.f * fma_(w, y, -x * z), -1.f, 1.f));
    rpy[2] = atan2_(2.f * fma_(x, y, w * z), fma_(-2.f, fma_(y, y, z * z), 1.f));
}

// scipy Rotation.from_euler('xyz', [r, p, y]).as_quat() -> wxyz (utils/state.py:59-60)
QS_HD void rpy_to_quat(const float rpy[3], float q[4]) {
    float sr, cr, sp, cp, sy, cy;
    sincos_(0.5f * rpy[0], &sr, &cr);
    sincos_(0.5f * rpy[1], &sp, &cp);
    sincos_(0.5f * rpy[2], &sy, &cy);
    q[0] = fma_(cr * cp, cy, sr * sp * sy);
    q[1] = fma_(sr * cp, cy, -cr * sp * sy);
    q[2] = fma_(cr * sp, cy, sr * cp * sy);
    q[3] = fma_(cr * cp, sy, -sr * sp * cy);
}

// ---------------------------------------------------------------------------------------
// observation
// ---------------------------------------------------------------------------------------
template <int MODE>
QS_HD void compute_obs(const QsParams& P, const Env& e, const float rpy[3], float* obs) {
    if constexpr (ModeTraits<MODE>::kGym) {
        // hover_env.py:126-136 + utils/normalization.py:17
        const float x[12] = {
            e.target[0] - e.b.p[0], e.target[1] - e.b.p[1], e.target[2] - e.b.p[2],
            rpy[0], rpy[1], rpy[2], e.b.v[0], e.b.v[1], e.b.v[2], e.b.w[0], e.b.w[1], e.b.w[2]};
#pragma unroll
        for (int i = 0; i < 12; ++i) obs[i] = fma_(x[i] - P.obs_lo[i], P.obs_scale[i], -1.0f);
    } else {
        // train_brax_ppo.py:366-368 (+ NaN -> 0 at :338 for the mjx brax env)
        const float x[21] = {
            e.b.p[0], e.b.p[1], e.b.p[2], e.b.q[0], e.b.q[1], e.b.q[2], e.b.q[3],
            e.b.th[0], e.b.th[1], e.b.th[2], e.b.th[3],
            e.b.v[0], e.b.v[1], e.b.v[2], e.b.w[0], e.b.w[1], e.b.w[2],
            e.b.s[0], e.b.s[1], e.b.s[2], e.b.s[3]};
#pragma unroll
        for (int i = 0; i < 21; ++i) {
            if constexpr (MODE == QS_MODE_MJX_BRAX) obs[i] = finite_(x[i]) ? x[i] : 0.0f;
            else obs[i] = x[i];
        }
    }
}

QS_HD bool body_finite(const Body& b) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 3; ++i) ok = ok && finite_(b.p[i]) && finite_(b.v[i]) && finite_(b.w[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) ok = ok && finite_(b.q[i]) && finite_(b.th[i]) && finite_(b.s[i]);
    return ok;
}

// ---------------------------------------------------------------------------------------
// reward / termination of the CURRENT state (after step_count has been advanced)
// ---------------------------------------------------------------------------------------
template <int MODE>
QS_HD void evaluate(const QsParams& P, const Tables& T, Env& e, const float* a, const float rpy[3], StepOut& o) {
    using M = ModeTraits<MODE>;
    float tgt[3];
    if constexpr (M::kTable) {
        // idx = min(step_count, N-1)   train_brax_ppo.py:319-321 / jax_mjx_quad_env.py:164
        const int idx = e.step_count < P.max_episode_steps - 1 ? e.step_count : P.max_episode_steps - 1;
        tgt[0] = T.target[3 * idx]; tgt[1] = T.target[3 * idx + 1]; tgt[2] = T.target[3 * idx + 2];
    } else if constexpr (MODE == QS_MODE_HOVER_BRAX) {
        tgt[0] = P.fixed_target[0]; tgt[1] = P.fixed_target[1]; tgt[2] = P.fixed_target[2];
    } else {
        tgt[0] = e.target[0]; tgt[1] = e.target[1]; tgt[2] = e.target[2];
    }
    const float dx = e.b.p[0] - tgt[0], dy = e.b.p[1] - tgt[1], dz = e.b.p[2] - tgt[2];
    const float e2 = fma_(dx, dx, fma_(dy, dy, dz * dz));
    float asq = 0.f;
    if (a) asq = fma_(a[0], a[0], fma_(a[1], a[1], fma_(a[2], a[2], a[3] * a[3])));
    o.truncated = 0.f;

    if constexpr (M::kGym) {
        // hover_env.py:138-157,188
        o.pos_error = sqrt_(e2);
        o.reward_hover = exp_(-e2);
        o.reward_action = 0.f;
        o.reward = o.reward_hover;
        const float s12[12] = {e.b.p[0], e.b.p[1], e.b.p[2], rpy[0], rpy[1], rpy[2],
                               e.b.v[0], e.b.v[1], e.b.v[2], e.b.w[0], e.b.w[1], e.b.w[2]};
        bool inside = true;
#pragma unroll
        for (int i = 0; i < 12; ++i)
            inside = inside && (s12[i] >= P.term_lo[i]) && (s12[i] <= P.term_hi[i]);   // false for NaN / +-Inf too
        o.done = inside ? 0.f : 1.f;
        o.truncated = (e.step_count >= P.max_episode_steps) ? 1.f : 0.f;
    } else if constexpr (MODE == QS_MODE_MJX_BRAX) {
        // train_brax_ppo.py:324-336
        const bool fin = body_finite(e.b);
        const bool out_xy = (fabsf(e.b.p[0]) > P.pos_limit_xy) || (fabsf(e.b.p[1]) > P.pos_limit_xy);
        const bool out_z = (e.b.p[2] < P.z_low) || (e.b.p[2] > P.z_high);
        const bool out_v = (fabsf(e.b.v[0]) > P.vel_limit) || (fabsf(e.b.v[1]) > P.vel_limit) ||
                           (fabsf(e.b.v[2]) > P.vel_limit);
        const bool valid = fin && !out_xy && !out_z && !out_v;
        const float raw = sqrt_(e2);
        o.pos_error = (valid && finite_(raw)) ? raw : 1e3f;
        o.reward_hover = exp_(-(o.pos_error * o.pos_error));
        o.reward_action = -P.action_penalty * asq;
        const float r = o.reward_hover + o.reward_action;
        o.reward = (valid && finite_(r)) ? r : -1.0f;
        o.done = valid ? 0.f : 1.f;
    } else if constexpr (MODE == QS_MODE_HOVER_BRAX) {
        // train_brax_ppo.py:143-159: exp(-2 e^2); the action term is reported but not added (Q2)
        o.pos_error = sqrt_(e2);
        o.reward_hover = exp_(-P.reward_k * (o.pos_error * o.pos_error));
        o.reward_action = -0.001f * asq;
        o.reward = o.reward_hover;
        const bool out_xy = (fabsf(e.b.p[0]) > P.pos_limit_xy) || (fabsf(e.b.p[1]) > P.pos_limit_xy);
        const bool out_z = (e.b.p[2] < P.z_low) || (e.b.p[2] > P.z_high);
        o.done = (out_xy || out_z) ? 1.f : 0.f;
    } else {
        // jax_mjx_quad_env.py:150,163-172: never terminates, truncates at max_episode_steps
        o.pos_error = sqrt_(e2);
        o.reward_hover = exp_(-(o.pos_error * o.pos_error));
        o.reward_action = 0.f;
        o.reward = o.reward_hover;
        o.done = 0.f;
        o.truncated = (e.step_count >= P.max_episode_steps) ? 1.f : 0.f;
    }
}

// ---------------------------------------------------------------------------------------
// reset: Philox(seed; global env id, episode, block, stream 0)
// ---------------------------------------------------------------------------------------
// rpy (gym modes): the Euler angles of the new attitude.  They are the drawn angles themselves, so the
// reset observation needs no quaternion -> Euler round trip (the reference's round trip through
// scipy changes them by < 1e-7 rad).
template <int MODE>
QS_HD void reset_env(const QsParams& P, const Tables& T, uint32_t gid, Env& e, float rpy[3]) {
    using M = ModeTraits<MODE>;
    rpy[0] = 0.f; rpy[1] = 0.f; rpy[2] = 0.f;
    e.step_count = 0;
    e.ep_steps = 0;
    e.done_prev = 0.f;
    e.voltage = P.v_nominal;
#pragma unroll
    for (int k = 0; k < 3; ++k) e.rate_int[k] = 0.f;       // RateControlWrapper.reset (rate_wrapper.py:108-111)
#pragma unroll
    for (int k = 0; k < 4; ++k) { e.b.th[k] = 0.f; e.b.s[k] = 0.f; e.prev_action[k] = 0.f; }
    if constexpr (M::kGym) {
        if (P.waypoint_mode) {
            // evaluate.py:487-497: start on waypoint 0, identity attitude, at rest; target = waypoint 1
            const int shape = (int)(gid % (uint32_t)P.wp_num_shapes);
            const int n = P.wp_count[shape];
            const double* wp = T.waypoints + (size_t)shape * QS_MAX_WP * 3;
            e.b.p[0] = (float)wp[0]; e.b.p[1] = (float)wp[1]; e.b.p[2] = (float)wp[2];
            e.b.q[0] = 1.f; e.b.q[1] = 0.f; e.b.q[2] = 0.f; e.b.q[3] = 0.f;
#pragma unroll
            for (int i = 0; i < 3; ++i) { e.b.v[i] = 0.f; e.b.w[i] = 0.f; }
            e.wp_idx = 1 % n;
            e.target[0] = (float)wp[3 * e.wp_idx]; e.target[1] = (float)wp[3 * e.wp_idx + 1];
            e.target[2] = (float)wp[3 * e.wp_idx + 2];
            return;
        }
        // hover_env.py:218-229: state12 ~ U(init_lo, init_hi) float32, target ~ U(target_lo, target_hi)
        float s12[12];
#pragma unroll
        for (int blk = 0; blk < 3; ++blk) {
            const U4 r = philox4x32_10(U4{gid, e.episode, (uint32_t)blk, STREAM_RESET}, P.philox_key);
            s12[4 * blk + 0] = uniform_(r.x, P.init_lo[4 * blk + 0], P.init_hi[4 * blk + 0]);
            s12[4 * blk + 1] = uniform_(r.y, P.init_lo[4 * blk + 1], P.init_hi[4 * blk + 1]);
            s12[4 * blk + 2] = uniform_(r.z, P.init_lo[4 * blk + 2], P.init_hi[4 * blk + 2]);
            s12[4 * blk + 3] = uniform_(r.w, P.init_lo[4 * blk + 3], P.init_hi[4 * blk + 3]);
        }
        e.b.p[0] = s12[0]; e.b.p[1] = s12[1]; e.b.p[2] = s12[2];
        rpy_to_quat(&s12[3], e.b.q);
        rpy[0] = s12[3]; rpy[1] = s12[4]; rpy[2] = s12[5];
        e.b.v[0] = s12[6]; e.b.v[1] = s12[7]; e.b.v[2] = s12[8];
        e.b.w[0] = s12[9]; e.b.w[1] = s12[10]; e.b.w[2] = s12[11];
        if constexpr (MODE == QS_MODE_HOVER_GYM) {
            const U4 r = philox4x32_10(U4{gid, e.episode, 3u, STREAM_RESET}, P.philox_key);
            e.target[0] = uniform_(r.x, P.target_lo[0], P.target_hi[0]);
            e.target[1] = uniform_(r.y, P.target_lo[1], P.target_hi[1]);
            e.target[2] = uniform_(r.z, P.target_lo[2], P.target_hi[2]);
        } else {
            // trajectory_follow_env.py:241-243 + Q7: the target stays at the spline's first point,
            // which is the start position
            e.target[0] = e.b.p[0]; e.target[1] = e.b.p[1]; e.target[2] = e.b.p[2];
        }
    } else if constexpr (MODE == QS_MODE_MJX_PLAYGROUND) {
        // jax_mjx_quad_env.py:114-122: make_data only -> qpos0, no noise
        e.b.p[0] = 0.f; e.b.p[1] = 0.f; e.b.p[2] = 0.f;
        e.b.q[0] = 1.f; e.b.q[1] = 0.f; e.b.q[2] = 0.f; e.b.q[3] = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) { e.b.v[i] = 0.f; e.b.w[i] = 0.f; }
    } else {
        // train_brax_ppo.py:244-271 / :102-118: nominal pose + U(-n, n) on every qpos and qvel entry
        float n21[24];
#pragma unroll
        for (int blk = 0; blk < 6; ++blk) {
            const U4 r = philox4x32_10(U4{gid, e.episode, (uint32_t)blk, STREAM_RESET}, P.philox_key);
            n21[4 * blk + 0] = uniform_(r.x, -P.reset_noise, P.reset_noise);
            n21[4 * blk + 1] = uniform_(r.y, -P.reset_noise, P.reset_noise);
            n21[4 * blk + 2] = uniform_(r.z, -P.reset_noise, P.reset_noise);
            n21[4 * blk + 3] = uniform_(r.w, -P.reset_noise, P.reset_noise);
        }
        e.b.p[0] = n21[0]; e.b.p[1] = n21[1]; e.b.p[2] = P.reset_z + n21[2];
        float q0 = 1.0f + n21[3], q1 = n21[4], q2 = n21[5], q3 = n21[6];
        if constexpr (MODE == QS_MODE_MJX_BRAX) {
            // quat / (||quat|| + 1e-8)   train_brax_ppo.py:265-268
            const float inv = 1.0f / (sqrt_(fma_(q0, q0, fma_(q1, q1, fma_(q2, q2, q3 * q3)))) + 1e-8f);
            q0 *= inv; q1 *= inv; q2 *= inv; q3 *= inv;
        }
        e.b.q[0] = q0; e.b.q[1] = q1; e.b.q[2] = q2; e.b.q[3] = q3;
#pragma unroll
        for (int k = 0; k < 4; ++k) e.b.th[k] = n21[7 + k];
        e.b.v[0] = n21[11]; e.b.v[1] = n21[12]; e.b.v[2] = n21[13];
        e.b.w[0] = n21[14]; e.b.w[1] = n21[15]; e.b.w[2] = n21[16];
#pragma unroll
        for (int k = 0; k < 4; ++k) e.b.s[k] = n21[17 + k];
    }
}

// ---------------------------------------------------------------------------------------
// waypoint advance (evaluate.py:540-557).  Returns true when the lap completed.
// Distances are float64 on the float32 state, exactly as the reference computes them.
// ---------------------------------------------------------------------------------------
QS_HD bool waypoint_advance(const QsParams& P, const Tables& T, uint32_t gid, Env& e) {
    const int shape = (int)(gid % (uint32_t)P.wp_num_shapes);
    const int n = P.wp_count[shape];
    const double* wp = T.waypoints + (size_t)shape * QS_MAX_WP * 3;
    const double dx = (double)e.b.p[0] - wp[3 * e.wp_idx];
    const double dy = (double)e.b.p[1] - wp[3 * e.wp_idx + 1];
    const double dz = (double)e.b.p[2] - wp[3 * e.wp_idx + 2];
#if defined(__CUDA_ARCH__)
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const double dist = __dsqrt_rn(d2);
#else
    volatile double xx = dx * dx, yy = dy * dy, zz = dz * dz;
    volatile double s1 = xx + yy;
    volatile double d2 = s1 + zz;
    const double dist = sqrt(d2);
#endif
    if (!(dist < (double)P.wp_reach_radius)) return false;
    e.wp_reached += 1;
    e.wp_idx = (e.wp_idx + 1) % n;
    if (e.wp_idx == 0) { e.laps += 1; return true; }
    e.target[0] = (float)wp[3 * e.wp_idx]; e.target[1] = (float)wp[3 * e.wp_idx + 1];
    e.target[2] = (float)wp[3 * e.wp_idx + 2];
    return false;
}

// ---------------------------------------------------------------------------------------
// one full env step incl. wrapper logic.  obs receives the observation the policy sees next
// (after auto-reset if one happened); term_obs (may be null) the pre-reset observation.
// first: pointer to this env's first_state column (plane stride `nenv`), RESTORE_FIRST only.
// ---------------------------------------------------------------------------------------
// DEFER_RESET: the Philox re-sampling of a finished gym env is NOT done here; the caller sees
// o.needs_reset and performs it (the kernels compact those lanes per block, see block_autoreset).
template <int MODE, bool DEFER_RESET = false, class F = FeatAll>
QS_HD void env_step(const QsParams& P, const Tables& T, uint32_t gid, Env& e, const float a[4],
                    float* obs, float* term_obs, const float* first, int nenv, StepOut& o) {
    using M = ModeTraits<MODE>;
    if constexpr (M::kBrax) {
        // AutoResetWrapper.step: steps <- 0 where the previous state was done
        if (P.auto_reset == QS_RESET_RESTORE_FIRST && e.done_prev != 0.f) e.ep_steps = 0;
    }
    float ctrl[4];
    if constexpr (M::kGym) {
        if (F::kRate && P.rate_wrapper) {
            // the policy commands body rates; the base env sees torques, _prev_action keeps the rate action
            float at[4];
            rate_to_torque(P, a, e.b.w, e.rate_int, at);
#pragma unroll
            for (int k = 0; k < 4; ++k) e.prev_action[k] = a[k];
            action_to_ctrl<F>(P, at, e.voltage, ctrl);
        } else {
            action_to_ctrl<F>(P, a, e.voltage, ctrl);
        }
    } else {
        action_to_ctrl<F>(P, a, e.voltage, ctrl);
    }
    physics_step(P, e.b, ctrl);
    e.step_count += 1;

    float rpy[3] = {0.f, 0.f, 0.f};
    if constexpr (M::kGym) quat_to_rpy(e.b.q, rpy);
    evaluate<MODE>(P, T, e, a, rpy, o);
    compute_obs<MODE>(P, e, rpy, obs);
    o.finished = false;
    o.needs_reset = false;

    if constexpr (M::kGym) {
        bool lap = false;
        if (F::kWaypoint && P.waypoint_mode) lap = waypoint_advance(P, T, gid, e);
        o.finished = (o.done != 0.f) || (o.truncated != 0.f) || lap;
        if (o.finished) {
            if (term_obs) {
#pragma unroll
                for (int i = 0; i < 12; ++i) term_obs[i] = obs[i];
            }
            if (P.auto_reset == QS_RESET_RESAMPLE) {
                // VecEnv semantics: the returned observation is the first one of the next episode
                e.episode += 1u;
                if constexpr (DEFER_RESET) {
                    o.needs_reset = true;
                } else {
                    reset_env<MODE>(P, T, gid, e, rpy);
                    compute_obs<MODE>(P, e, rpy, obs);
                }
            }
        }
    } else if constexpr (M::kBrax) {
        if (P.episode_length > 0) {
            // EpisodeWrapper.step (action_repeat = 1)
            e.ep_steps += 1;
            const bool over = e.ep_steps >= P.episode_length;
            o.truncated = over ? 1.0f - o.done : 0.0f;
            o.done = over ? 1.0f : o.done;
        }
        o.finished = o.done != 0.f;
        if (o.finished && term_obs) {
#pragma unroll
            for (int i = 0; i < 21; ++i) term_obs[i] = obs[i];
        }
        if (P.auto_reset == QS_RESET_RESTORE_FIRST) {
            if (o.finished) {
                // AutoResetWrapper: pipeline_state, obs <- first_*; step_count keeps counting (Q5)
                float* dst[21] = {&e.b.p[0], &e.b.p[1], &e.b.p[2], &e.b.q[0], &e.b.q[1], &e.b.q[2], &e.b.q[3],
                                  &e.b.th[0], &e.b.th[1], &e.b.th[2], &e.b.th[3],
                                  &e.b.v[0], &e.b.v[1], &e.b.v[2], &e.b.w[0], &e.b.w[1], &e.b.w[2],
                                  &e.b.s[0], &e.b.s[1], &e.b.s[2], &e.b.s[3]};
#pragma unroll
                for (int i = 0; i < 21; ++i) *dst[i] = first[(size_t)i * nenv];
                compute_obs<MODE>(P, e, rpy, obs);
            }
            e.done_prev = o.done;
        }
    } else {
        o.finished = (o.truncated != 0.f);
        if (o.finished && term_obs) {
#pragma unroll
            for (int i = 0; i < 21; ++i) term_obs[i] = obs[i];
        }
    }
}

}  // namespace qs
