// TEST INFRASTRUCTURE -- not product code, never shipped, never loaded by the package.
//
// Compiles the very same per-env source the sm_100a kernels are built from
// (csrc/qs_env.cuh, qs_dynamics.cuh, qs_philox.cuh) with plain g++ and runs it in a loop
// over envs.  The dev box has no GPU; this lets the closed-form dynamics and the env
// semantics be checked against oracle/ before any GPU time is spent.  GPU parity tests
// (-m gpu) call the real library through the C ABI instead.
#include <stdint.h>
#include <stddef.h>
#include <math.h>

#include "../../uav_reinforcement_learning_control_b200/csrc/qs_env.cuh"
#include "../../uav_reinforcement_learning_control_b200/csrc/qs_traj.cuh"
#include "../../uav_reinforcement_learning_control_b200/csrc/qs_umma_desc.cuh"

using namespace qs;

template <int MODE>
static void step_all(const QsParams& P, const Tables& T, int n, float* state, const float* action, float* obs,
                     float* reward, float* done, float* trunc, float* metrics, float* term_obs, const float* first) {
    constexpr int D = ModeTraits<MODE>::kObsDim;
    for (int i = 0; i < n; ++i) {
        Env e;
        load_env<MODE>(P, state, n, i, e);
        StepOut so;
        float o_[D], tobs[D];
        env_step<MODE>(P, T, P.env_id_offset + (uint32_t)i, e, action + 4 * (size_t)i, o_, term_obs ? tobs : nullptr,
                       first ? first + i : nullptr, n, so);
        store_env<MODE>(P, state, n, i, e);
        for (int k = 0; k < D; ++k) obs[(size_t)i * D + k] = o_[k];
        reward[i] = so.reward; done[i] = so.done;
        if (trunc) trunc[i] = so.truncated;
        if (metrics) {
            metrics[i] = so.pos_error; metrics[(size_t)n + i] = so.reward_hover;
            metrics[2 * (size_t)n + i] = so.reward_action; metrics[3 * (size_t)n + i] = so.reward;
        }
        if (term_obs && so.finished) for (int k = 0; k < D; ++k) term_obs[(size_t)i * D + k] = tobs[k];
    }
}

template <int MODE>
static void reset_all(const QsParams& P, const Tables& T, int n, float* state, const uint8_t* mask, float* obs,
                      float* first) {
    constexpr int D = ModeTraits<MODE>::kObsDim;
    for (int i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        Env e;
        load_env<MODE>(P, state, n, i, e, true);
        float rpy0[3];
        reset_env<MODE>(P, T, P.env_id_offset + (uint32_t)i, e, rpy0);
        store_env<MODE>(P, state, n, i, e);
        if (first) {
            const float x[21] = {e.b.p[0], e.b.p[1], e.b.p[2], e.b.q[0], e.b.q[1], e.b.q[2], e.b.q[3],
                                 e.b.th[0], e.b.th[1], e.b.th[2], e.b.th[3], e.b.v[0], e.b.v[1], e.b.v[2],
                                 e.b.w[0], e.b.w[1], e.b.w[2], e.b.s[0], e.b.s[1], e.b.s[2], e.b.s[3]};
            for (int k = 0; k < 21; ++k) first[(size_t)k * n + i] = x[k];
        }
        if (obs) {
            float rpy[3] = {0.f, 0.f, 0.f};
            if (ModeTraits<MODE>::kGym) quat_to_rpy(e.b.q, rpy);
            float o_[D];
            compute_obs<MODE>(P, e, rpy, o_);
            for (int k = 0; k < D; ++k) obs[(size_t)i * D + k] = o_[k];
        }
    }
}

template <int MODE>
static void observe_all(const QsParams& P, const Tables& T, int n, const float* state, const float* action,
                        float* obs, float* reward, float* done) {
    constexpr int D = ModeTraits<MODE>::kObsDim;
    for (int i = 0; i < n; ++i) {
        Env e;
        load_env<MODE>(P, state, n, i, e);
        float rpy[3] = {0.f, 0.f, 0.f};
        if (ModeTraits<MODE>::kGym) quat_to_rpy(e.b.q, rpy);
        StepOut so;
        evaluate<MODE>(P, T, e, action ? action + 4 * (size_t)i : nullptr, rpy, so);
        float o_[D];
        compute_obs<MODE>(P, e, rpy, o_);
        if (obs) for (int k = 0; k < D; ++k) obs[(size_t)i * D + k] = o_[k];
        if (reward) reward[i] = so.reward;
        if (done) done[i] = so.done;
    }
}

#define HH_DISPATCH(CALL)                                                                    \
    switch (P->mode) {                                                                       \
        case QS_MODE_MJX_BRAX: { constexpr int M_ = QS_MODE_MJX_BRAX; CALL; break; }         \
        case QS_MODE_HOVER_GYM: { constexpr int M_ = QS_MODE_HOVER_GYM; CALL; break; }       \
        case QS_MODE_TRAJ_GYM: { constexpr int M_ = QS_MODE_TRAJ_GYM; CALL; break; }         \
        case QS_MODE_HOVER_BRAX: { constexpr int M_ = QS_MODE_HOVER_BRAX; CALL; break; }     \
        case QS_MODE_MJX_PLAYGROUND: { constexpr int M_ = QS_MODE_MJX_PLAYGROUND; CALL; break; } \
        default: return -1;                                                                  \
    }

extern "C" {

int hh_params_size(void) { return (int)sizeof(QsParams); }

int hh_step(const QsParams* P, const float* target_table, const double* waypoints, int n, float* state,
            const float* action, float* obs, float* reward, float* done, float* trunc, float* metrics,
            float* term_obs, const float* first) {
    Tables T{target_table, waypoints};
    HH_DISPATCH((step_all<M_>(*P, T, n, state, action, obs, reward, done, trunc, metrics, term_obs, first)));
    return 0;
}

int hh_reset(const QsParams* P, const float* target_table, const double* waypoints, int n, float* state,
             const uint8_t* mask, float* obs, float* first) {
    Tables T{target_table, waypoints};
    HH_DISPATCH((reset_all<M_>(*P, T, n, state, mask, obs, first)));
    return 0;
}

int hh_observe(const QsParams* P, const float* target_table, const double* waypoints, int n, const float* state,
               const float* action, float* obs, float* reward, float* done) {
    Tables T{target_table, waypoints};
    HH_DISPATCH((observe_all<M_>(*P, T, n, state, action, obs, reward, done)));
    return 0;
}

int hh_physics(const QsParams* P, int n, float* state, const float* ctrl) {
    for (int i = 0; i < n; ++i) {
        Env e;
        load_env<QS_MODE_HOVER_BRAX>(*P, state, n, i, e);
        physics_step(*P, e.b, ctrl + 4 * (size_t)i);
        store_env<QS_MODE_HOVER_BRAX>(*P, state, n, i, e);
    }
    return 0;
}

int hh_traj_info(const QsParams* P, int n, const uint32_t* episode, const int32_t* sample_index, float* out9) {
    for (int i = 0; i < n; ++i) traj_info_eval(*P, P->env_id_offset + (uint32_t)i, episode[i], sample_index[i], out9 + 9 * (size_t)i);
    return 0;
}

void hh_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
    U4 r = philox4x32_10(U4{c0, c1, c2, c3}, k0, k1);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// the tcgen05 descriptor packing the tensor-core kernels use (qs_umma_desc.cuh), for tests/test_umma_desc.py
uint64_t hh_umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) { return tc::make_desc(saddr, lbo_bytes, sbo_bytes); }
uint32_t hh_umma_idesc(int M, int N, int a_mn, int b_mn) { return tc::idesc_mn(M, N, a_mn, b_mn); }

}  // extern "C"
