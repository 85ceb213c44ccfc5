// qs_dynamics.cuh -- closed-form restatement of one mjx.step / mj_step for the quad model.
//
// Replaces: mjx.step(model, data) at train_brax_ppo.py:317 / jax_mjx_quad_env.py:144 and
// mujoco.mj_step at envs/hover_env.py:180 for model/drone/drone.xml (free base + 4 passive
// z-axis rotors + site-transmission motors + inertia-box fluid drag, Euler integrator).
//
// Generalised coordinates   q = [p_O (world), quat wxyz, theta_1..4]
// generalised velocities    v = [dp_O/dt (world axes), omega (BODY axes), s_1..4]
//
// Because the rotors are balanced and axisymmetric about their hinge (checked by the
// loader), the composite COM c, the inertia about it I_C and the rotor axial inertia J are
// constant in the base frame, and MuJoCo's 10x10 system M qacc = passive + actuator - bias
// reduces to Newton-Euler for a gyrostat about the composite COM:
//
//   I_eff * domega = tau_C - omega x (I_C omega + z h) - z * sum_k rho_k tau'_k
//   ds_k           = (tau'_k - J_k domega_z) / Js_k
//   a_b            = f_b / m - omega x (omega x c) - domega x c          (base-frame comps)
//   dv_O           = R a_b
// with h = sum J_k s_k, tau'_k = hinge torque on rotor k (fluid axial torque - damping),
// Js_k = J_k + armature_k + dt*damping_k (MuJoCo's implicit-in-velocity Euler damping),
// rho_k = J_k / Js_k, I_eff = I_C - z z^T sum J_k^2 / Js_k, f_b / tau_C = total external
// force / torque about the COM in base-frame components (thrust sites, gravity, fluid).
// Then semi-implicit Euler: v += dt*qacc; positions advance with the NEW velocities; the
// quaternion is advanced by the body-frame angular velocity (mju_quatIntegrate) and
// renormalised.  DESIGN.md derives this from the generic pipeline in oracle/.
#pragma once

#include "qs_math.cuh"
#include "../../include/quadsim_abi.h"

#ifndef QS_ROTOR_UNROLL
#define QS_ROTOR_UNROLL 4      /* unroll factor of the four-rotor drag loop; 1 or 2 cost 19 % of the step kernel's speed:
                                   the four independent rotors are the ILP that hides the FP32 latency at 7 warps/scheduler */
#endif
#define QS_PRAGMA_(x) _Pragma(#x)
#define QS_UNROLL_(n) QS_PRAGMA_(unroll n)

namespace qs {

// R = float: one env per thread (every kernel but one, and the host test harness); R = f2 (qs_pack2.cuh): the same
// quantity of TWO envs per thread on the packed FP32 pipe (qs_step2.cuh).  One source serves both.
template <class R>
struct BodyT {
    R p[3];    // base origin, world
    R q[4];    // quaternion w x y z (body -> world)
    R th[4];   // rotor angles
    R v[3];    // origin velocity, world axes
    R w[3];    // angular velocity, body axes
    R s[4];    // rotor rates
};
using Body = BodyT<float>;

// rotation matrix of a unit quaternion, row-major r[3*i+j]
template <class R>
QS_HD void quat_to_mat(const R q[4], R r[9]) {
    const R w = q[0], x = q[1], y = q[2], z = q[3];
    const R xx = x * x, yy = y * y, zz = z * z;
    const R xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
    r[0] = fma_(-2.f, yy + zz, 1.f); r[1] = 2.f * (xy - wz);           r[2] = 2.f * (xz + wy);
    r[3] = 2.f * (xy + wz);           r[4] = fma_(-2.f, xx + zz, 1.f); r[5] = 2.f * (yz - wx);
    r[6] = 2.f * (xz - wy);           r[7] = 2.f * (yz + wx);           r[8] = fma_(-2.f, xx + yy, 1.f);
}

// one drag component: -(visc + quad*|u|) * u
template <class R>
QS_HD R drag_(float visc, float quad, R u) { return -fma_(quad, abs_(u), visc) * u; }

// quaternion integration coefficients for the body rate with squared norm n2: cos(|w| dt / 2) and sin(|w| dt / 2) / |w|;
// |w| < 1e-15 (mjMINVAL): mju_quatIntegrate applies no rotation
QS_HD void quat_step_coeffs(float n2, float dt, float* cs, float* k) {
    if (n2 > 1e-30f) {
        const float inv = rsqrt_(n2);
        const float n = n2 * inv;
        float sn;
        sincos_(0.5f * dt * n, &sn, cs);
        *k = sn * inv;
    } else {
        *cs = 1.f; *k = 0.f;
    }
}
#if defined(__CUDACC__)
__device__ __forceinline__ void quat_step_coeffs(f2 n2, float dt, f2* cs, f2* k) {
    const f2 inv = rsqrt_(n2);                    // a zero lane gives inf / NaN here and is replaced below
    const f2 n = n2 * inv;
    f2 sn, c;
    sincos_((0.5f * dt) * n, &sn, &c);
    const b2 rot = gt_(n2, 1e-30f);
    *cs = sel_(rot, c, 1.f);
    *k = sel_(rot, sn * inv, 0.f);
}
#endif

// Advance `b` by one timestep under motor forces `ctrl` (N, clamped to ctrlrange here as
// MuJoCo's fwd_actuation does).
template <class T>
QS_HD void physics_step(const QsParams& P, BodyT<T>& b, const T ctrl[4]) {
    // ---- attitude (quaternion is used normalised, as in mj_kinematics) ----------------
    T qn[4];
    {
        const T n2 = fma_(b.q[0], b.q[0], fma_(b.q[1], b.q[1], fma_(b.q[2], b.q[2], b.q[3] * b.q[3])));
        const T inv = rsqrt_(n2);
        qn[0] = b.q[0] * inv; qn[1] = b.q[1] * inv; qn[2] = b.q[2] * inv; qn[3] = b.q[3] * inv;
    }
    T R[9];
    quat_to_mat(qn, R);
    const T wx = b.w[0], wy = b.w[1], wz = b.w[2];
    // origin velocity in base-frame components: vb = R^T v
    const T vbx = fma_(R[0], b.v[0], fma_(R[3], b.v[1], R[6] * b.v[2]));
    const T vby = fma_(R[1], b.v[0], fma_(R[4], b.v[1], R[7] * b.v[2]));
    const T vbz = fma_(R[2], b.v[0], fma_(R[5], b.v[1], R[8] * b.v[2]));

    // ---- motors: site wrench, torque referred to the composite COM --------------------
    T F[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) F[k] = clamp_(ctrl[k], P.ctrl_lo[k], P.ctrl_hi[k]);
    T f[3], t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        f[i] = fma_(P.wrench[4 * i + 0], F[0], fma_(P.wrench[4 * i + 1], F[1],
               fma_(P.wrench[4 * i + 2], F[2], P.wrench[4 * i + 3] * F[3])));
        t[i] = fma_(P.wrench[12 + 4 * i + 0], F[0], fma_(P.wrench[12 + 4 * i + 1], F[1],
               fma_(P.wrench[12 + 4 * i + 2], F[2], P.wrench[12 + 4 * i + 3] * F[3])));
    }
    // gravity m*g along world z, in base components = m*gz * (third row of R); no torque about the COM
    {
        const float mg = P.mass * P.gz;
        f[0] = fma_(mg, R[6], f[0]); f[1] = fma_(mg, R[7], f[1]); f[2] = fma_(mg, R[8], f[2]);
    }

    // ---- fluid, base body (inertial frame == base frame, COM at the origin) -----------
    {
        const T fx = drag_(P.base_lin_visc, P.base_lin_quad[0], vbx);
        const T fy = drag_(P.base_lin_visc, P.base_lin_quad[1], vby);
        const T fz = drag_(P.base_lin_visc, P.base_lin_quad[2], vbz);
        f[0] += fx; f[1] += fy; f[2] += fz;
        // applied at the origin: lever arm about the composite COM is -c
        t[0] += drag_(P.base_ang_visc, P.base_ang_quad[0], wx) - (P.com[1] * fz - P.com[2] * fy);
        t[1] += drag_(P.base_ang_visc, P.base_ang_quad[1], wy) - (P.com[2] * fx - P.com[0] * fz);
        t[2] += drag_(P.base_ang_visc, P.base_ang_quad[2], wz) - (P.com[0] * fy - P.com[1] * fx);
    }

    // ---- fluid, rotors: the box frame spins with theta_k --------------------------------
    T th_[4];   // hinge torque tau'_k
    QS_UNROLL_(QS_ROTOR_UNROLL)
    for (int k = 0; k < 4; ++k) {
        T sn, cs;
        sincos_fast_(b.th[k], &sn, &cs);
        const float rx = P.rotor_r[3 * k], ry = P.rotor_r[3 * k + 1], rz = P.rotor_r[3 * k + 2];
        // velocity of the rotor COM, base components: vb + w x r
        const T ux = vbx + (wy * rz - wz * ry);
        const T uy = vby + (wz * rx - wx * rz);
        const T uz = vbz + (wx * ry - wy * rx);
        const T u1 = fma_(cs, ux, sn * uy), u2 = fma_(cs, uy, -sn * ux);
        const T f1 = drag_(P.rot_lin_visc[k], P.rot_lin_quad_lat[k], u1);
        const T f2_ = drag_(P.rot_lin_visc[k], P.rot_lin_quad_lat[k], u2);
        const T fz = drag_(P.rot_lin_visc[k], P.rot_lin_quad_ax[k], uz);
        const T fx = fma_(cs, f1, -sn * f2_), fy = fma_(sn, f1, cs * f2_);
        // rotor angular velocity, base components: w + s_k z
        const T w1 = fma_(cs, wx, sn * wy), w2 = fma_(cs, wy, -sn * wx), wa = wz + b.s[k];
        const T t1 = drag_(P.rot_ang_visc[k], P.rot_ang_quad_lat[k], w1);
        const T t2 = drag_(P.rot_ang_visc[k], P.rot_ang_quad_lat[k], w2);
        const T ta = drag_(P.rot_ang_visc[k], P.rot_ang_quad_ax[k], wa);
        const T tx = fma_(cs, t1, -sn * t2), ty = fma_(sn, t1, cs * t2);
        const float dx = P.rotor_d[3 * k], dy = P.rotor_d[3 * k + 1], dz = P.rotor_d[3 * k + 2];
        f[0] += fx; f[1] += fy; f[2] += fz;
        t[0] += tx + (dy * fz - dz * fy);
        t[1] += ty + (dz * fx - dx * fz);
        t[2] += ta + (dx * fy - dy * fx);
        th_[k] = fma_(-P.rotor_damp[k], b.s[k], ta);
    }

    // ---- gyrostat Euler equation about the COM ------------------------------------------
    const T h = fma_(P.rotor_J[0], b.s[0], fma_(P.rotor_J[1], b.s[1], fma_(P.rotor_J[2], b.s[2], P.rotor_J[3] * b.s[3])));
    const T Lx = fma_(P.I_C[0], wx, fma_(P.I_C[1], wy, P.I_C[2] * wz));
    const T Ly = fma_(P.I_C[3], wx, fma_(P.I_C[4], wy, P.I_C[5] * wz));
    const T Lz = fma_(P.I_C[6], wx, fma_(P.I_C[7], wy, P.I_C[8] * wz)) + h;
    t[0] -= wy * Lz - wz * Ly;
    t[1] -= wz * Lx - wx * Lz;
    t[2] -= wx * Ly - wy * Lx;
    t[2] -= fma_(P.rotor_rho[0], th_[0], fma_(P.rotor_rho[1], th_[1], fma_(P.rotor_rho[2], th_[2], P.rotor_rho[3] * th_[3])));
    const T dwx = fma_(P.Ieff_inv[0], t[0], fma_(P.Ieff_inv[1], t[1], P.Ieff_inv[2] * t[2]));
    const T dwy = fma_(P.Ieff_inv[3], t[0], fma_(P.Ieff_inv[4], t[1], P.Ieff_inv[5] * t[2]));
    const T dwz = fma_(P.Ieff_inv[6], t[0], fma_(P.Ieff_inv[7], t[1], P.Ieff_inv[8] * t[2]));

    // ---- origin acceleration, base components --------------------------------------------
    const float cx = P.com[0], cy = P.com[1], cz = P.com[2];
    const T wc = fma_(wx, cx, fma_(wy, cy, wz * cz));
    const T w2n = fma_(wx, wx, fma_(wy, wy, wz * wz));
    // w x (w x c) = w (w.c) - c |w|^2 ;  dw x c
    const T ax = fma_(f[0], P.inv_mass, -(fma_(wx, wc, -cx * w2n)) - (dwy * cz - dwz * cy));
    const T ay = fma_(f[1], P.inv_mass, -(fma_(wy, wc, -cy * w2n)) - (dwz * cx - dwx * cz));
    const T az = fma_(f[2], P.inv_mass, -(fma_(wz, wc, -cz * w2n)) - (dwx * cy - dwy * cx));

    // ---- semi-implicit Euler -------------------------------------------------------------
    const float dt = P.dt;
    b.v[0] = fma_(dt, fma_(R[0], ax, fma_(R[1], ay, R[2] * az)), b.v[0]);
    b.v[1] = fma_(dt, fma_(R[3], ax, fma_(R[4], ay, R[5] * az)), b.v[1]);
    b.v[2] = fma_(dt, fma_(R[6], ax, fma_(R[7], ay, R[8] * az)), b.v[2]);
    b.w[0] = fma_(dt, dwx, wx); b.w[1] = fma_(dt, dwy, wy); b.w[2] = fma_(dt, dwz, wz);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const T ds = fma_(-P.rotor_J[k], dwz, th_[k]) * P.rotor_invJs[k];
        b.s[k] = fma_(dt, ds, b.s[k]);
        b.th[k] = fma_(dt, b.s[k], b.th[k]);
    }
    b.p[0] = fma_(dt, b.v[0], b.p[0]); b.p[1] = fma_(dt, b.v[1], b.p[1]); b.p[2] = fma_(dt, b.v[2], b.p[2]);

    // quaternion: q <- normalize(qn * [cos(|w|dt/2), sin(|w|dt/2) w/|w|]) with the NEW w
    {
        const T nx = b.w[0], ny = b.w[1], nz = b.w[2];
        const T n2 = fma_(nx, nx, fma_(ny, ny, nz * nz));
        T cs, k;
        quat_step_coeffs(n2, dt, &cs, &k);
        const T ex = k * nx, ey = k * ny, ez = k * nz;
        const T w0 = qn[0], x0 = qn[1], y0 = qn[2], z0 = qn[3];
        T rw = fma_(w0, cs, -fma_(x0, ex, fma_(y0, ey, z0 * ez)));
        T rx = fma_(w0, ex, fma_(x0, cs, fma_(y0, ez, -z0 * ey)));
        T ry = fma_(w0, ey, fma_(y0, cs, fma_(z0, ex, -x0 * ez)));
        T rz = fma_(w0, ez, fma_(z0, cs, fma_(x0, ey, -y0 * ex)));
        const T m2 = fma_(rw, rw, fma_(rx, rx, fma_(ry, ry, rz * rz)));
        const T minv = rsqrt_(m2);
        b.q[0] = rw * minv; b.q[1] = rx * minv; b.q[2] = ry * minv; b.q[3] = rz * minv;
    }
}

}  // namespace qs
