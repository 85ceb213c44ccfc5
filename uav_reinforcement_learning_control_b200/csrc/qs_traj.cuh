// qs_traj.cuh -- TrajectoryFollowEnv's spline reference trajectory, evaluated on demand.
//
// Replaces: TrajectoryFollowEnv._sample_sinusoid_trajectory (envs/trajectory_follow_env.py:176-218: 3..5 random
// waypoints per axis through scipy CubicSpline(bc_type='natural'), sampled into [2048][3] pos / vel / acc arrays at
// every reset) and the info["target" | "target_vel" | "target_acc"] look-ups (:162-168, :245-250).
//
// The reference stores 3 x 2048 x 3 floats per env and fits three splines on the CPU at every reset.  Here nothing is
// stored: the spline of an episode is a pure function of (seed, global env id, episode index) -- the same Philox
// draws that produced the start position plus one extra stream for the centre / waypoint count / offsets -- so the
// info kernel re-derives the <= 5 knots, solves the <= 3 x 3 tridiagonal system for the knot second derivatives and
// evaluates position / velocity / acceleration at the requested sample, all in float64 like scipy, cast to float32
// like the reference's arrays.  oracle/traj_spline.py states the draw layout and checks this against scipy itself.
#pragma once

#include "qs_env.cuh"

namespace qs {

enum : uint32_t { STREAM_TRAJ = 3u };

// natural cubic spline through y[0..n-1] at knots j*h (n in 3..5), second-derivative form; value and first two
// derivatives at t in [0, (n-1) h]
QS_HD void natural_spline_eval(const double* y, int n, double h, double t, double* s, double* s1, double* s2) {
    double m[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    {
        // Thomas elimination of  m[i-1] + 4 m[i] + m[i+1] = 6 (y[i-1] - 2 y[i] + y[i+1]) / h^2,  m[0] = m[n-1] = 0
        const int k = n - 2;
        double c[3], r[3];
        const double ih2 = 6.0 / (h * h);
        c[0] = 0.25;
        r[0] = 0.25 * ih2 * (y[0] - 2.0 * y[1] + y[2]);
        for (int i = 1; i < k; ++i) {
            const double den = 1.0 / (4.0 - c[i - 1]);
            c[i] = den;
            r[i] = (ih2 * (y[i] - 2.0 * y[i + 1] + y[i + 2]) - r[i - 1]) * den;
        }
        m[k] = r[k - 1];
        for (int i = k - 2; i >= 0; --i) m[i + 1] = r[i] - c[i] * m[i + 2];
    }
    int j = (int)(t / h);
    j = j < n - 2 ? j : n - 2;
    const double a = (double)(j + 1) * h - t, b = t - (double)j * h;
    const double mj = m[j], mk = m[j + 1], yj = y[j], yk = y[j + 1];
    *s = (mj * a * a * a + mk * b * b * b) / (6.0 * h) + (yj - mj * h * h / 6.0) * a / h + (yk - mk * h * h / 6.0) * b / h;
    *s1 = (-mj * a * a + mk * b * b) / (2.0 * h) + (yk - yj) / h - (mk - mj) * h / 6.0;
    *s2 = (mj * a + mk * b) / h;
}

// out9 = target(3) | target_vel(3) | target_acc(3) of env `gid` in episode `episode` at sample index `idx`
QS_HD void traj_info_eval(const QsParams& P, uint32_t gid, uint32_t episode, int idx, float out9[9]) {
    const int N = P.max_episode_steps;
    idx = idx < 0 ? 0 : (idx > N - 1 ? N - 1 : idx);
    // time base: np.linspace(0, duration, N) or arange(N) * dt  (trajectory_follow_env.py:187-191)
    double T, t;
    if (P.spline_duration > 0.f) {
        T = (double)P.spline_duration;
        t = idx == N - 1 ? T : (double)idx * (T / (double)(N - 1));
    } else {
        T = (double)(N - 1) * (double)P.dt;
        t = (double)idx * (double)P.dt;
    }
    const U4 r0 = philox4x32_10(U4{gid, episode, 0u, STREAM_RESET}, P.philox_key);
    const float start[3] = {uniform_(r0.x, P.init_lo[0], P.init_hi[0]), uniform_(r0.y, P.init_lo[1], P.init_hi[1]),
                            uniform_(r0.z, P.init_lo[2], P.init_hi[2])};
    const U4 rc = philox4x32_10(U4{gid, episode, 0u, STREAM_TRAJ}, P.philox_key);
    const float centre[3] = {uniform_(rc.x, P.traj_center_lo[0], P.traj_center_hi[0]),
                             uniform_(rc.y, P.traj_center_lo[1], P.traj_center_hi[1]),
                             uniform_(rc.z, P.traj_center_lo[2], P.traj_center_hi[2])};
    int extra = (int)(u01_(rc.w) * 3.0f);
    extra = extra > 2 ? 2 : extra;
    const int n = 3 + extra;                                   // integers(3, 6)
    uint32_t w[16];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const U4 r = philox4x32_10(U4{gid, episode, (uint32_t)(b + 1), STREAM_TRAJ}, P.philox_key);
        w[4 * b] = r.x; w[4 * b + 1] = r.y; w[4 * b + 2] = r.z; w[4 * b + 3] = r.w;
    }
    const double h = T / (double)(n - 1);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double y[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) y[j] = (double)centre[a] + (double)uniform_(w[5 * a + j], -P.traj_amp[a], P.traj_amp[a]);
        y[0] = (double)start[a];                               // the spline starts at the drone (:208-209)
        double s, s1, s2;
        natural_spline_eval(y, n, h, t, &s, &s1, &s2);
        out9[a] = (float)s; out9[3 + a] = (float)s1; out9[6 + a] = (float)s2;
    }
}

}  // namespace qs
