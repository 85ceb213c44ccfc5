/*
 * quadsim_abi.h -- C ABI of libquadsim (sm_100a), the drop-in boundary of the
 * batched quadrotor env step / rollout / GAE hot path.
 *
 * The reference has no FFI for this path: the boundary is two Python duck-typed
 * protocols (Brax Env: train_brax_ppo.py:232-356; Gymnasium Env:
 * envs/hover_env.py:159-238).  Each entry point below states which reference
 * interface it replaces.  All pointers are DEVICE pointers unless the name ends in
 * _host; all buffers are caller-owned; `stream` is a cudaStream_t passed as void*
 * (NULL = legacy default stream).  Every call returns 0 on success or a negative
 * QS_E* code, never throws, never falls back to the CPU.  One handle per
 * (device, env shard); a handle is not thread-safe, distinct handles are independent.
 *
 * Data layout.  Env state is planar struct-of-arrays, float32/int32, `QS_NPLANES`
 * planes of `num_envs` elements each: plane p of env i is state[p * num_envs + i].
 *   0..10  qpos  (x y z | qw qx qy qz | rotor angles 1..4)      reference: data.qpos
 *   11..20 qvel  (vx vy vz world | wx wy wz body | rotor rates)  reference: data.qvel
 *   21..23 target position                                       hover_env.py:227
 *   24     step_count (int32)            train_brax_ppo.py:319 / hover_env.py:182
 *   25     battery voltage                                       hover_env.py:102-109
 *   26     episode index (uint32, Philox counter word)
 *   27     ep_steps (int32, Brax EpisodeWrapper info["steps"])
 *   28     waypoint index (int32)                                evaluate.py:496,551
 *   29     waypoints reached (int32)                             evaluate.py:550
 *   30     laps completed (int32)                                evaluate.py:553
 *   31     done flag of the previous step (float32 0/1; Brax AutoResetWrapper)
 *   32..34 rate-wrapper integral state (N m)                    envs/rate_wrapper.py:64
 *   35..38 previous policy action (RelPosActWrapper obs)        envs/hover_env.py:165, rate_wrapper.py:100-104
 *   39     reserved
 * Actions are [num_envs][4] float32 (row-major, 16 B per env); observations are
 * [num_envs][obs_dim] float32 row-major; reward/done/truncated are [num_envs] float32.
 */
#ifndef QUADSIM_ABI_H
#define QUADSIM_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QS_ABI_VERSION 1
#define QS_NPLANES 40
#define QS_NQ 11
#define QS_NV 10
#define QS_MAX_WP 64
#define QS_MAX_SHAPES 4

/* error codes */
#define QS_OK 0
#define QS_EINVAL (-1)     /* bad argument / inconsistent config            */
#define QS_ECUDA (-2)      /* CUDA runtime error, see qs_last_error_string  */
#define QS_ENOMEM (-3)
#define QS_EUNSUPPORTED (-4)

/* env semantics selector (SURVEY 8a "Env variants") */
enum QsMode {
    QS_MODE_MJX_BRAX = 0,       /* JaxMJXQuadBraxEnv     train_brax_ppo.py:179-368 */
    QS_MODE_HOVER_GYM = 1,      /* HoverEnv              envs/hover_env.py          */
    QS_MODE_TRAJ_GYM = 2,       /* TrajectoryFollowEnv   envs/trajectory_follow_env.py */
    QS_MODE_HOVER_BRAX = 3,     /* QuadHoverBraxEnv      train_brax_ppo.py:39-176   */
    QS_MODE_MJX_PLAYGROUND = 4  /* JaxMJXQuadEnv         envs/jax_mjx_quad_env.py   */
};

enum QsAutoReset {
    QS_RESET_NONE = 0,
    QS_RESET_RESTORE_FIRST = 1, /* Brax AutoResetWrapper: restore the episode-0 state */
    QS_RESET_RESAMPLE = 2       /* VecEnv semantics: fresh Philox sample per episode  */
};

/*
 * Plain-old-data parameter block: model constants (derived in float64 from the MJCF
 * by the host, stored as float32) + env semantics.  It is passed BY VALUE to every
 * kernel, i.e. it lives in the constant bank for the lifetime of the launch.
 */
typedef struct QsParams {
    /* --- integrator / composite body ------------------------------------ */
    float dt, gz, mass, inv_mass;
    float com[3];            /* composite COM in the base frame                  */
    float I_C[9];            /* composite inertia about the COM (row-major)       */
    float Ieff_inv[9];       /* inv(I_C - zz^T sum J^2/Js)                       */
    float rotor_J[4], rotor_invJs[4], rotor_rho[4] /* J/Js */, rotor_damp[4];
    float rotor_r[12];       /* rotor COM in base frame, [k][3]                   */
    float rotor_d[12];       /* rotor COM relative to composite COM, [k][3]       */
    float wrench[24];        /* [6][4]: body force ; torque about COM, per N      */
    float ctrl_lo[4], ctrl_hi[4];
    /* --- fluid ------------------------------------------------------------ */
    float base_lin_visc, base_ang_visc;
    float base_lin_quad[3], base_ang_quad[3];
    float rot_lin_visc[4], rot_lin_quad_ax[4], rot_lin_quad_lat[4];
    float rot_ang_visc[4], rot_ang_quad_ax[4], rot_ang_quad_lat[4];
    /* --- action path (utils/drone_config.py, hover_env.py:60-65,94-100) ---- */
    float act_lo[4], act_hi[4];       /* physical action bounds                 */
    float mix_inv[16];                /* A^-1, row-major                          */
    float max_motor_thrust;
    int32_t pre_clip_action;          /* Q1: Brax variants clip the denormalised action */
    /* --- battery (hover_env.py:102-109) ------------------------------------ */
    int32_t battery;
    float v_nominal, v_min, v_drop_base, v_drop_load;
    /* --- observation / reward / termination -------------------------------- */
    int32_t mode, obs_dim;
    float obs_lo[12], obs_scale[12];  /* obs = (x - lo) * scale - 1, scale = 2/(hi-lo) */
    float term_lo[12], term_hi[12];   /* gym modes: inclusive box on the 12-D state */
    float pos_limit_xy, z_low, z_high, vel_limit;   /* brax modes                 */
    float reward_k;                   /* exp(-k e^2): 1, or 2 for hover_brax (Q2)  */
    float action_penalty;             /* 0.001 for mjx_brax, else 0               */
    float fixed_target[3];            /* hover_brax (0,0,h)                       */
    /* --- episode handling -------------------------------------------------- */
    int32_t max_episode_steps;        /* gym truncation / length of the mjx target table */
    int32_t episode_length;           /* Brax EpisodeWrapper; 0 = wrapper off     */
    int32_t auto_reset;               /* enum QsAutoReset                         */
    /* --- reset distribution ------------------------------------------------ */
    float init_lo[12], init_hi[12];   /* hover_env.py:42-45                       */
    float target_lo[3], target_hi[3]; /* hover_env.py:48-51                       */
    float reset_noise;                /* brax modes: U(-n, n) on all qpos/qvel    */
    float reset_z;                    /* brax modes: nominal height               */
    uint32_t seed_lo, seed_hi;
    uint32_t env_id_offset;           /* global id of local env 0 (sharding)      */
    uint32_t philox_key[20];          /* Philox4x32-10 round keys of (seed_lo, seed_hi): [2r] = seed_lo + r*0x9E3779B9,
                                         [2r+1] = seed_hi + r*0xBB67AE85 -- precomputed so the kernels need no key schedule */
    /* --- waypoint tracking (evaluate.py:440-557) --------------------------- */
    int32_t waypoint_mode;            /* 0 off */
    int32_t wp_num_shapes;
    int32_t wp_count[QS_MAX_SHAPES];
    float wp_reach_radius;
    /* --- rate-control action wrapper (envs/rate_wrapper.py:16-111) ------------ */
    int32_t rate_wrapper;             /* 1: action = [thrust, roll/pitch/yaw RATE] -> torques by a PI loop */
    float rate_max;                   /* rad/s for |a| = 1 (360 deg/s)            */
    float rate_kd[3];                 /* P gains (inertia-scaled)                 */
    float rate_inertia[3];            /* IXX, IYY, IZZ of utils/drone_config.py   */
    float rate_ki, rate_imax;         /* integral gain (torque space), anti-windup clamp */
    float max_torque;                 /* normalisation of the torque command (0.5 N m) */
    /* --- TrajectoryFollowEnv spline reference (envs/trajectory_follow_env.py:55-58,176-218) --- */
    float spline_duration;            /* traj_duration_seconds (30); <= 0: sample times are arange(N) * dt */
    float traj_center_lo[3], traj_center_hi[3];   /* centre ~ U(lo, hi)                         */
    float traj_amp[3];                /* waypoint offsets ~ U(-amp, amp): 0.6, 0.6, 0.4          */
    int32_t reserved[5];
} QsParams;

typedef struct QsEngine* QsHandle;

/* library identity / diagnostics */
int qs_abi_version(void);
const char* qs_last_error_string(void);
/* sizeof(QsParams) as compiled, so a binding can verify its mirror of the struct */
int qs_params_size(void);

/*
 * qs_create: replaces the env constructors (train_brax_ppo.py:180-230,
 * envs/hover_env.py:15-100): binds a parameter block to a device and `num_envs`.
 *   target_table_host : [max_episode_steps][3] float32 target positions for the mjx
 *                       modes (train_brax_ppo.py:358-364), may be NULL otherwise.
 *   waypoints_host    : [wp_num_shapes][QS_MAX_WP][3] float64 (utils/trajectories.py),
 *                       may be NULL when waypoint_mode == 0.
 */
int qs_create(const QsParams* params, int32_t num_envs, int32_t device,
              const float* target_table_host, const double* waypoints_host, QsHandle* out);
int qs_destroy(QsHandle h);
int qs_num_envs(QsHandle h);
int qs_get_params(QsHandle h, QsParams* out);

/*
 * qs_reset: replaces Env.reset (train_brax_ppo.py:244-289; hover_env.py:200-238).
 * Re-initialises every env whose mask entry is non-zero (mask == NULL: all) with the
 * Philox draw for (seed, global env id, episode index); writes obs for those envs.
 * first_state (optional, [21][num_envs]) receives the qpos/qvel of the new episode
 * (Brax AutoResetWrapper info["first_pipeline_state"]).
 */
int qs_reset(QsHandle h, float* state, const uint8_t* mask, float* obs, float* first_state, void* stream);

/*
 * qs_step: replaces Env.step (train_brax_ppo.py:307-356; hover_env.py:159-198) fused
 * with the vectorising wrappers' episode logic (Brax Episode/AutoReset wrappers; SB3
 * VecEnv auto-reset) as configured in QsParams.  In place on `state`.
 *   metrics      : optional [4][num_envs]: pos_error, reward_hover, reward_action, reward
 *   truncated    : optional [num_envs]
 *   terminal_obs : optional [num_envs][obs_dim], written only for envs that finished
 *   first_state  : required when auto_reset == QS_RESET_RESTORE_FIRST
 */
int qs_step(QsHandle h, float* state, const float* action, float* obs, float* reward, float* done,
            float* truncated, float* metrics, float* terminal_obs, const float* first_state, void* stream);

/*
 * qs_observe: obs / reward / done of the CURRENT state without stepping
 * (hover_env.py:126-157 _get_obs/_get_reward/_is_terminated; train_brax_ppo.py:324-338).
 * `action` may be NULL (action penalty = 0).  Used for injected-state parity.
 */
int qs_observe(QsHandle h, const float* state, const float* action, float* obs, float* reward,
               float* done, void* stream);

/*
 * qs_physics_step: the bare mjx.step / mj_step replacement (train_brax_ppo.py:317,
 * hover_env.py:180): ctrl [num_envs][4] motor forces in N -> advances planes 0..20 only.
 */
int qs_physics_step(QsHandle h, float* state, const float* ctrl, void* stream);

/*
 * qs_rollout_random: state-resident dynamics-only rollout, T env steps per launch with
 * in-kernel Philox actions U(-1,1)^4 (stream 2), auto-reset per QsParams.  Accumulates
 * per-env sums: stats [4][num_envs] += (sum reward, episodes finished, obs checksum, steps).
 * `t0` is the global step index of the first step (Philox counter word).
 */
int qs_rollout_random(QsHandle h, float* state, int32_t T, uint32_t t0, float* stats, const float* first_state,
                      void* stream);

/*
 * Policy rollout (replaces brax acting.generate_unroll / SB3 collect_rollouts around the
 * env, train_brax_ppo.py:589-620, train.py:133-137).  See qs_policy.h section below.
 */
typedef struct QsPolicyDesc {
    int32_t obs_dim;        /* 12 or 21 */
    int32_t hidden;         /* 128 */
    int32_t act_dim;        /* 4 */
    int32_t dist;           /* 0 = SB3 diagonal Gaussian with state-independent log_std,
                               1 = Brax tanh-normal, actor head emits loc|raw_scale (8) */
    int32_t deterministic;  /* 1: action = mean (tanh(mean) for dist 1) */
    float bootstrap_gamma;  /* > 0: SB3 timeout bootstrap, reward += gamma * V(terminal_obs) on truncation */
    int32_t tensor_cores;   /* 0: fp32 FMA path; 1: tcgen05/TMEM path (bf16 operands, fp32 accumulate): hover_gym dist 0|1,
                               traj_gym dist 0, mjx_brax / hover_brax dist 1 */
    int32_t sample_seed;    /* qs_ppo_grad, dist 1 only: key of the entropy-term noise (brax draws a fresh action sample per loss
                               evaluation); the caller advances it every update.  Ignored elsewhere. */
} QsPolicyDesc;

/* number of float32 in the packed parameter vector for a policy description:
 * actor W1[obs][H] b1[H] W2[H][H] b2[H] W3[H][A'] b3[A'] (A' = act_dim or 2*act_dim),
 * critic W1[obs][H] b1[H] W2[H][H] b2[H] W3[H] b3[1], log_std[act_dim] (dist 0 only),
 * obs_mean[obs] obs_inv_std[obs].  Weights are [in][out] row-major. */
int qs_policy_param_count(const QsPolicyDesc* desc);

/*
 * qs_rollout_policy: T steps of {policy forward -> sample -> env step} with env state
 * resident on chip.  Trajectory buffers are time-major:
 *   traj_obs [T][B][obs_dim], traj_act [T][B][4], traj_logp/value/reward/done/trunc [T][B];
 *   last_value [B] = V(obs after the last step); last_obs [B][obs_dim] (optional) that obs.
 * traj_act holds the RAW Gaussian sample (pre-clip for dist 0, pre-tanh for dist 1), as SB3's
 * RolloutBuffer and brax's policy_extras["raw_action"] do.
 * Any trajectory pointer may be NULL to skip that stream.
 */
int qs_rollout_policy(QsHandle h, float* state, const QsPolicyDesc* desc, const float* policy_params,
                      int32_t T, uint32_t t0, float* last_obs,
                      float* traj_obs, float* traj_act, float* traj_logp, float* traj_value,
                      float* traj_reward, float* traj_done, float* traj_trunc,
                      float* last_value, const float* first_state, void* stream);

/*
 * qs_gae: reverse-time scan over [T][B] (SB3 RolloutBuffer.compute_returns_and_advantage
 * semantics when brax_form == 0; brax compute_gae semantics when brax_form == 1).
 */
int qs_gae(int32_t T, int32_t B, const float* reward, const float* value, const float* done,
           const float* trunc, const float* last_value, float gamma, float lam, int32_t brax_form,
           float* adv, float* ret, void* stream);

/*
 * PPO update on device.  Replaces the caller of the rollout, SB3 ``PPO.train`` as configured by
 * train.py:50-68 (MlpPolicy pi=[128,128] vf=[128,128] ReLU, Adam eps 1e-5, clip_range, ent_coef, vf_coef, max_grad_norm,
 * per-minibatch advantage normalisation), and the loss of brax ``ppo_train.train`` (train_brax_ppo.py:589-620: tanh-normal
 * policy, 21-D observation) on precomputed advantages.  Covers obs_dim 12 | 21 with dist 0 | 1; handle-free like qs_gae,
 * runs on the current device.  All pointers are device memory except the scalars.
 * normalize_adv: 0 off, 1 (adv - mean) / (unbiased std + 1e-8) as SB3 / torch, 2 the same with the population std as brax / jnp.
 * dist 1: loss = -mean(min(rho A, clip(rho) A)) + vf_coef mean((ret - V)^2) - ent_coef mean(H) with H the tanh-normal entropy
 * estimate brax uses (one fresh sample per row, keyed by desc->sample_seed); brax's value loss is vf_coef = 0.25.
 *
 * qs_ppo_grad: gradient of   -mean(min(r A, clip(r, 1-c, 1+c) A)) + vf_coef mean((ret - V)^2) - ent_coef H(pi)
 * over the minibatch rows idx[0..n) (idx == NULL: rows 0..n-1) of the rollout buffers obs [N][12], act [N][4] (raw
 * Gaussian samples), old_logp / adv / ret [N].  Forward and backward GEMMs run on tcgen05 (bf16 operands, fp32 TMEM
 * accumulation).  grad receives qs_policy_param_count floats in the packed layout (zeros for obs_mean / obs_inv_std)
 * followed by 8 statistics: sum policy loss, sum value loss, clipped samples, sum approx-KL, samples, 0, 0, 0.
 * workspace: qs_ppo_workspace_bytes(desc) bytes, zero-filled once by the caller before the first call.
 *
 * qs_ppo_adam: params -= Adam(clip_by_global_norm(grad * grad_scale, max_grad_norm)) on the trained part of the packed
 * vector (torch.optim.Adam and torch.nn.utils.clip_grad_norm_ semantics; max_grad_norm <= 0: no clipping).  `step` is
 * the 1-based optimiser step; grad_scale = 1 / world_size after a sum all-reduce.  norm_out (optional): pre-clip norm.
 *
 * qs_ppo_permutation: out[0..n) (device, int32) = a pseudo-random permutation of 0..n-1 keyed by (seed, epoch) -- the
 * per-epoch shuffle of SB3's RolloutBuffer.get (np.random.permutation) as a keyed bijection (4-round Feistel network +
 * cycle walking) evaluated independently per element; slices of it are the `idx` minibatches of qs_ppo_grad.
 */
int64_t qs_ppo_workspace_bytes(const QsPolicyDesc* desc);
int qs_ppo_permutation(int32_t n, uint64_t seed, uint32_t epoch, int32_t* out, void* stream);
int qs_ppo_grad(const QsPolicyDesc* desc, const float* policy_params, const float* obs, const float* act,
                const float* old_logp, const float* adv, const float* ret, const int32_t* idx, int32_t n,
                float clip_range, float vf_coef, float ent_coef, int32_t normalize_adv, void* workspace, float* grad,
                void* stream);
int qs_ppo_adam(const QsPolicyDesc* desc, float* policy_params, const float* grad, float* m, float* v, int32_t step,
                float lr, float beta1, float beta2, float eps, float max_grad_norm, float grad_scale, float* norm_out,
                void* stream);

/*
 * Packed sample rows.  qs_ppo_pack turns the five time-major rollout / GAE arrays into ONE 128-byte row per sample,
 * packed[N][32] = obs[obs_dim] | act[4] | old_logp | adv | ret | 0.., (128-byte aligned), once per rollout; qs_ppo_grad_packed
 * is qs_ppo_grad reading its minibatch rows from there: a random row is then one full cache line instead of 4-7 partial
 * sectors of five arrays (SB3's RolloutBuffer.get gathers the same five arrays, stable-baselines3 [third party]).  `adv` is the
 * contiguous advantage array again: the minibatch statistics (normalize_adv != 0) gather 4-byte values, which is cheaper
 * from a dense array than from 128-byte rows.
 */
int qs_ppo_pack(const QsPolicyDesc* desc, const float* obs, const float* act, const float* old_logp, const float* adv,
                const float* ret, int64_t n, float* packed, void* stream);
int qs_ppo_grad_packed(const QsPolicyDesc* desc, const float* policy_params, const float* packed, const float* adv,
                       const int32_t* idx, int32_t n, float clip_range, float vf_coef, float ent_coef, int32_t normalize_adv,
                       void* workspace, float* grad, void* stream);

/*
 * Running observation normaliser of the Brax trainer (normalize_observations=True, train_brax_ppo.py:611;
 * brax.training.acme.running_statistics): merges obs [n][obs_dim] into running = {count, mean[obs_dim], M2[obs_dim]}
 * (device doubles, zero before the first call) with the parallel Welford update and writes mean_out / inv_std_out
 * (normally pointers INTO the packed parameter vector: its obs_mean / obs_inv_std entries), inv_std = 1 / clip(sqrt(M2 /
 * count), std_min, std_max).  workspace: qs_obs_stats_workspace_bytes(obs_dim) bytes of device scratch.
 */
int64_t qs_obs_stats_workspace_bytes(int32_t obs_dim);
int qs_obs_stats_update(const float* obs, int64_t n, int32_t obs_dim, double* running, float* mean_out, float* inv_std_out,
                        float std_min, float std_max, void* workspace, void* stream);

/*
 * Multi-GPU form of the update's gradient exchange (one process per GPU; what brax's ``lax.pmean`` of the gradients does
 * inside ``ppo_train.train``, train_brax_ppo.py:589-620, and what a data-parallel SB3 learner would need): instead of
 * reduce -> NCCL all-reduce -> Adam, every rank leaves its gradient in a slot of a CUDA-IPC-exported buffer and ONE
 * kernel per rank waits for the peers' slots (flags in peer memory), sums them over NVLink in rank order, clips by the
 * global norm and applies Adam -- the parameters stay bitwise identical on all ranks by construction.
 *   qs_ppo_comm_create  : allocates this rank's buffer (world <= 8)
 *   qs_ppo_comm_export  : writes the 64-byte IPC handle of it (exchange it with any host-side all-gather)
 *   qs_ppo_comm_import  : maps rank `peer`'s buffer from its handle
 *   qs_ppo_comm_slot    : device pointer to pass as `grad` to qs_ppo_grad for update number `epoch` (1, 2, 3, ...;
 *                         the same count on every rank; slots alternate by parity)
 *   qs_ppo_adam_peer    : the fused wait + sum + clip + Adam for that `epoch`; stats_acc (optional, device, 8 floats)
 *                         += the world-summed loss statistics.  A peer that does not arrive within QS_PEER_TIMEOUT_MS
 *                         (environment, default 20 000 ms) makes the kernel raise an error word and return WITHOUT
 *                         touching the parameters (no trap, the context survives).
 *   qs_ppo_comm_error   : synchronous read of that error word: QS_OK, or QS_ECUDA with the late rank in the message
 *   tear-down           : every rank qs_ppo_comm_close_peers, a host-side barrier, then qs_ppo_comm_destroy (an exported
 *                         buffer must not be freed while a peer still maps it)
 */
typedef struct QsPpoComm QsPpoComm;
int qs_ppo_comm_create(const QsPolicyDesc* desc, int32_t world, int32_t rank, QsPpoComm** out);
int qs_ppo_comm_export(QsPpoComm* comm, void* handle64);
int qs_ppo_comm_import(QsPpoComm* comm, int32_t peer, const void* handle64);
void* qs_ppo_comm_slot(QsPpoComm* comm, uint32_t epoch);
int qs_ppo_adam_peer(const QsPolicyDesc* desc, QsPpoComm* comm, uint32_t epoch, float* policy_params, float* m, float* v,
                     int32_t step, float lr, float beta1, float beta2, float eps, float max_grad_norm, float* norm_out,
                     float* stats_acc, void* stream);
int qs_ppo_comm_error(QsPpoComm* comm);
int qs_ppo_comm_close_peers(QsPpoComm* comm);   /* unmap the peers' buffers; call on every rank (then synchronise the ranks) before any rank destroys */
int qs_ppo_comm_destroy(QsPpoComm* comm);

/*
 * qs_ppo_update_epoch: ONE epoch of minibatch updates -- what SB3's PPO.train() (train.py:50-68) / brax's sgd_step scan
 * (train_brax_ppo.py:589-620) run per pass over the rollout -- launched back to back from native code: for minibatch
 * k = 0 .. num_minibatches-1 over rows perm[k mb .. (k+1) mb) (mb = n_total / num_minibatches; the last one also takes the
 * remainder) of the packed sample rows: advantage statistics, qs_ppo_grad_packed, then qs_ppo_adam with step step0 + k
 * (comm == NULL; `grad` is the P + 8 float scratch) or, with a peer communicator, the gradient goes to slot epoch0 + k and
 * qs_ppo_adam_peer sums the ranks.  stats_acc (optional, device, 8 floats) += every minibatch's statistics.  desc's
 * sample_seed is replaced by (sample_seed0 + k) & 0x7fffffff per minibatch.  Bitwise identical to issuing the calls one by one.
 */
int qs_ppo_update_epoch(const QsPolicyDesc* desc, float* policy_params, const float* packed, const float* adv,
                        const int32_t* perm, int32_t n_total, int32_t num_minibatches, float clip_range, float vf_coef,
                        float ent_coef, int32_t normalize_adv, float* m, float* v, int32_t step0, float lr, float beta1,
                        float beta2, float eps, float max_grad_norm, QsPpoComm* comm, uint32_t epoch0, void* workspace,
                        float* grad, float* stats_acc, float* norm_out, uint64_t sample_seed0, void* stream);

/*
 * qs_traj_info: TrajectoryFollowEnv's info["target" | "target_vel" | "target_acc"] (envs/trajectory_follow_env.py:
 * 162-168 in step, :245-250 in reset), out9 [B][9] float32 = pos(3) | vel(3) | acc(3) of each env's natural-cubic-
 * spline reference (:176-218).  Nothing is stored per env: the spline of an episode is re-derived from the Philox
 * draws of (seed, global env id, episode).  `episode` / `sample_index` (device, [B]; either may be NULL) override
 * the state's episode plane (26) and the default index clamp(step_count - 1, 0, N - 1): a caller that wants the
 * info of the step that just ran -- including envs that were auto-reset by it -- passes copies of planes 26 and 24
 * taken before the step.  QS_MODE_TRAJ_GYM only.
 */
int qs_traj_info(QsHandle h, const float* state, const uint32_t* episode, const int32_t* sample_index, float* out9,
                 void* stream);

/*
 * Host-buffer convenience entry (what a non-torch binding of Env.step would call):
 * copies action_host H2D, steps, copies obs/reward/done D2H, synchronises.
 */
int qs_step_host(QsHandle h, float* state, const float* action_host, float* obs_host, float* reward_host,
                 float* done_host, void* stream);
/* the same, additionally returning the `truncated` flags (hover_env.py:188; Gymnasium's fifth return value) when
 * trunc_host is not NULL.  All host buffers must be page-locked (QS_EINVAL otherwise: pageable memory would silently
 * serialise the copy / compute overlap this call is built on). */
int qs_step_host_ex(QsHandle h, float* state, const float* action_host, float* obs_host, float* reward_host,
                    float* done_host, float* trunc_host, void* stream);

/* the same with the flags as bytes (0 / 1) -- the dtype the reference returns: Gymnasium's `terminated` / `truncated` are
 * bools (hover_env.py:186-198), SB3's `dones` a bool array -- which also takes 3 of the 56 bytes per env-step off the
 * device-to-host link this call is bound by.  num_envs must be a multiple of 4. */
int qs_step_host_bytes(QsHandle h, float* state, const float* action_host, float* obs_host, float* reward_host,
                    uint8_t* done_host, uint8_t* trunc_host, void* stream);

/* launch accounting for bench.py's gpu_launches claim */
uint64_t qs_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* QUADSIM_ABI_H */
