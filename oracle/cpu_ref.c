/*
 * ORACLE (test infrastructure, not product code) -- PARITY UNPINNED for the physics.
 *
 * Plain-C float64 restatement of the same generic pipeline as oracle/mujoco_pipeline.py
 * (what mujoco.mj_step / mjx.step compute for the reference's model; reference call sites
 * train_brax_ppo.py:317, envs/hover_env.py:180; model model/drone/drone.xml), plus the
 * HoverEnv step/reset semantics around it (envs/hover_env.py:159-238) so that the
 * reference's CPU-runnable configuration ("16 envs, random actions, 512-step episodes")
 * and bounded samples of the 1M-env workload can be TIMED on the host cores as
 * bench.py's cpu_baseline / `--impl reference` arm (kind "port": MuJoCo itself cannot be
 * installed here).  It does strictly less work than the reference's CPU path (no collision
 * broad-phase, no constraint solver, no Python), so it is an optimistic CPU baseline.
 *
 * It is tree-driven (knows nothing about quadrotors); tests check it against the NumPy
 * restatement to 1e-12 and both against the analytic known answers.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>

#define MAXB 8
#define MAXV 16
#define MAXU 8
#define JNT_FREE 0
#define JNT_HINGE 3

typedef struct {
    int nb, njnt, nq, nv, nu, nsite;
    int parent[MAXB];
    double pos[MAXB][3], quat[MAXB][4], ipos[MAXB][3], iquat[MAXB][4], mass[MAXB], inertia[MAXB][3];
    int jtype[MAXB], jbody[MAXB], qadr[MAXB], vadr[MAXB];
    double jpos[MAXB][3], jaxis[MAXB][3], damping[MAXV], armature[MAXV];
    int site_body[MAXB];
    double site_pos[MAXB][3], site_quat[MAXB][4];
    int act_site[MAXU], ctrl_limited[MAXU];
    double gear[MAXU][6], ctrl_lo[MAXU], ctrl_hi[MAXU];
    double dt, gravity[3], density, viscosity;
    double box[MAXB][3];
} OrcModel;

/* ---- small vector helpers ------------------------------------------------------------ */
static void cross3(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static void qmul(const double a[4], const double b[4], double o[4]) {
    o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}
static void q2m(const double q[4], double R[9]) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
static void mv(const double R[9], const double v[3], double o[3]) {
    o[0] = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    o[1] = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    o[2] = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
}
static void mtv(const double R[9], const double v[3], double o[3]) {
    o[0] = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
    o[1] = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
    o[2] = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
}
static void qnorm(double q[4]) {
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

typedef struct {
    int kind[MAXV];            /* 0 lin, 1 rot */
    double axis[MAXV][3], anchor[MAXV][3];
    int affects[MAXB][MAXV];
} Dofs;

/* Jacobian columns of a world point attached to body b */
static void jac(const OrcModel* m, const Dofs* d, int b, const double P[3], double Jv[3][MAXV], double Jw[3][MAXV]) {
    for (int v = 0; v < m->nv; ++v) {
        for (int k = 0; k < 3; ++k) { Jv[k][v] = 0; Jw[k][v] = 0; }
        if (!d->affects[b][v]) continue;
        if (d->kind[v] == 0) {
            for (int k = 0; k < 3; ++k) Jv[k][v] = d->axis[v][k];
        } else {
            double r[3] = {P[0] - d->anchor[v][0], P[1] - d->anchor[v][1], P[2] - d->anchor[v][2]}, c[3];
            cross3(d->axis[v], r, c);
            for (int k = 0; k < 3; ++k) { Jw[k][v] = d->axis[v][k]; Jv[k][v] = c[k]; }
        }
    }
}

/* One mj_step / mjx.step for one env.  qpos[nq], qvel[nv], ctrl[nu] in place. */
void orc_step_one(const OrcModel* m, double* qpos, double* qvel, const double* ctrl) {
    const int nb = m->nb, nv = m->nv;
    double xpos[MAXB][3], xquat[MAXB][4], R[MAXB][9];
    Dofs d;
    memset(&d, 0, sizeof(d));
    xpos[0][0] = xpos[0][1] = xpos[0][2] = 0; xquat[0][0] = 1; xquat[0][1] = xquat[0][2] = xquat[0][3] = 0;
    q2m(xquat[0], R[0]);
    double omega[MAXB][3] = {{0}}, alpha[MAXB][3] = {{0}}, vorg[MAXB][3] = {{0}}, aorg[MAXB][3] = {{0}};
    for (int b = 1; b < nb; ++b) {
        const int p = m->parent[b];
        double t[3];
        mv(R[p], m->pos[b], t);
        double pos[3] = {xpos[p][0] + t[0], xpos[p][1] + t[1], xpos[p][2] + t[2]};
        double quat[4];
        qmul(xquat[p], m->quat[b], quat);
        for (int v = 0; v < nv; ++v) d.affects[b][v] = d.affects[p][v];
        /* parent's motion carried to this origin */
        double dd[3] = {pos[0] - xpos[p][0], pos[1] - xpos[p][1], pos[2] - xpos[p][2]}, c1[3], c2[3];
        double w[3] = {omega[p][0], omega[p][1], omega[p][2]}, al[3] = {alpha[p][0], alpha[p][1], alpha[p][2]};
        double vv[3], aa[3];
        cross3(omega[p], dd, c1);
        for (int k = 0; k < 3; ++k) vv[k] = vorg[p][k] + c1[k];
        cross3(omega[p], c1, c2); cross3(alpha[p], dd, c1);
        for (int k = 0; k < 3; ++k) aa[k] = aorg[p][k] + c1[k] + c2[k];
        for (int j = 0; j < m->njnt; ++j) {
            if (m->jbody[j] != b) continue;
            const int qa = m->qadr[j], va = m->vadr[j];
            if (m->jtype[j] == JNT_FREE) {
                for (int k = 0; k < 3; ++k) pos[k] = qpos[qa + k];
                for (int k = 0; k < 4; ++k) quat[k] = qpos[qa + 3 + k];
                qnorm(quat);
                double Rb[9];
                q2m(quat, Rb);
                for (int k = 0; k < 3; ++k) {
                    d.kind[va + k] = 0; d.axis[va + k][0] = d.axis[va + k][1] = d.axis[va + k][2] = 0; d.axis[va + k][k] = 1;
                    d.affects[b][va + k] = 1;
                    d.kind[va + 3 + k] = 1;
                    for (int r = 0; r < 3; ++r) { d.axis[va + 3 + k][r] = Rb[3 * r + k]; d.anchor[va + 3 + k][r] = pos[r]; }
                    d.affects[b][va + 3 + k] = 1;
                }
                for (int k = 0; k < 3; ++k) vv[k] = qvel[va + k];
                mv(Rb, &qvel[va + 3], w);
                for (int k = 0; k < 3; ++k) { aa[k] = 0; al[k] = 0; }
            } else {
                double Rb[9], anchor[3], axl[3], axw[3], dq[4], q2[4], R2[9];
                q2m(quat, Rb);
                mv(Rb, m->jpos[j], t);
                for (int k = 0; k < 3; ++k) anchor[k] = pos[k] + t[k];
                const double an = sqrt(m->jaxis[j][0] * m->jaxis[j][0] + m->jaxis[j][1] * m->jaxis[j][1] + m->jaxis[j][2] * m->jaxis[j][2]);
                for (int k = 0; k < 3; ++k) axl[k] = m->jaxis[j][k] / an;
                mv(Rb, axl, axw);
                const double h = 0.5 * qpos[qa];
                dq[0] = cos(h); dq[1] = sin(h) * axl[0]; dq[2] = sin(h) * axl[1]; dq[3] = sin(h) * axl[2];
                qmul(quat, dq, q2);
                for (int k = 0; k < 4; ++k) quat[k] = q2[k];
                q2m(quat, R2);
                mv(R2, m->jpos[j], t);
                for (int k = 0; k < 3; ++k) pos[k] = anchor[k] - t[k];
                d.kind[va] = 1; d.affects[b][va] = 1;
                for (int k = 0; k < 3; ++k) { d.axis[va][k] = axw[k]; d.anchor[va][k] = anchor[k]; }
                /* velocities: anchor moves with the parent side */
                const double s = qvel[va];
                double da[3] = {anchor[0] - xpos[p][0], anchor[1] - xpos[p][1], anchor[2] - xpos[p][2]};
                double va_[3], aa_[3], wn[3], r[3] = {pos[0] - anchor[0], pos[1] - anchor[1], pos[2] - anchor[2]};
                cross3(omega[p], da, c1);
                for (int k = 0; k < 3; ++k) va_[k] = vorg[p][k] + c1[k];
                cross3(omega[p], c1, c2); cross3(alpha[p], da, c1);
                for (int k = 0; k < 3; ++k) aa_[k] = aorg[p][k] + c1[k] + c2[k];
                double sax[3] = {s * axw[0], s * axw[1], s * axw[2]};
                for (int k = 0; k < 3; ++k) wn[k] = w[k] + sax[k];
                cross3(w, sax, c1);
                for (int k = 0; k < 3; ++k) al[k] += c1[k];
                cross3(wn, r, c1);
                for (int k = 0; k < 3; ++k) vv[k] = va_[k] + c1[k];
                cross3(wn, c1, c2); cross3(al, r, c1);
                for (int k = 0; k < 3; ++k) { aa[k] = aa_[k] + c1[k] + c2[k]; w[k] = wn[k]; }
            }
        }
        for (int k = 0; k < 3; ++k) { xpos[b][k] = pos[k]; omega[b][k] = w[k]; alpha[b][k] = al[k]; vorg[b][k] = vv[k]; aorg[b][k] = aa[k]; }
        for (int k = 0; k < 4; ++k) xquat[b][k] = quat[k];
        q2m(quat, R[b]);
    }

    double M[MAXV][MAXV], rhs[MAXV];
    memset(M, 0, sizeof(M));
    for (int v = 0; v < nv; ++v) { M[v][v] = m->armature[v]; rhs[v] = -m->damping[v] * qvel[v]; }
    double Jv[3][MAXV], Jw[3][MAXV];
    for (int b = 1; b < nb; ++b) {
        const double mb = m->mass[b];
        if (mb <= 0) continue;
        double t[3], xipos[3], iq[4], Xi[9], Iw[9], tmp[9];
        mv(R[b], m->ipos[b], t);
        for (int k = 0; k < 3; ++k) xipos[k] = xpos[b][k] + t[k];
        qmul(xquat[b], m->iquat[b], iq);
        q2m(iq, Xi);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) tmp[3 * r + c] = Xi[3 * r + c] * m->inertia[b][c];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
            Iw[3 * r + c] = tmp[3 * r] * Xi[3 * c] + tmp[3 * r + 1] * Xi[3 * c + 1] + tmp[3 * r + 2] * Xi[3 * c + 2];
        jac(m, &d, b, xipos, Jv, Jw);
        for (int i = 0; i < nv; ++i) {
            double IJ[3];
            for (int r = 0; r < 3; ++r) IJ[r] = Iw[3 * r] * Jw[0][i] + Iw[3 * r + 1] * Jw[1][i] + Iw[3 * r + 2] * Jw[2][i];
            for (int j = 0; j < nv; ++j)
                M[j][i] += mb * (Jv[0][j] * Jv[0][i] + Jv[1][j] * Jv[1][i] + Jv[2][j] * Jv[2][i]) +
                           Jw[0][j] * IJ[0] + Jw[1][j] * IJ[1] + Jw[2][j] * IJ[2];
        }
        double rc[3] = {xipos[0] - xpos[b][0], xipos[1] - xpos[b][1], xipos[2] - xpos[b][2]}, c1[3], c2[3], c3[3];
        cross3(omega[b], rc, c1); cross3(omega[b], c1, c2); cross3(alpha[b], rc, c3);
        double f[3], nn[3], Iwv[3], Ial[3];
        for (int k = 0; k < 3; ++k) f[k] = mb * (aorg[b][k] + c3[k] + c2[k] - m->gravity[k]);
        mv(Iw, omega[b], Iwv); mv(Iw, alpha[b], Ial);
        cross3(omega[b], Iwv, nn);
        for (int k = 0; k < 3; ++k) nn[k] += Ial[k];
        for (int v = 0; v < nv; ++v)
            rhs[v] -= Jv[0][v] * f[0] + Jv[1][v] * f[1] + Jv[2][v] * f[2] + Jw[0][v] * nn[0] + Jw[1][v] * nn[1] + Jw[2][v] * nn[2];
        if (m->density > 0 || m->viscosity > 0) {
            double vcom[3], lw[3], lv[3], lt[3] = {0, 0, 0}, lf[3] = {0, 0, 0}, fw[3], tw[3];
            for (int k = 0; k < 3; ++k) vcom[k] = vorg[b][k] + c1[k];
            mtv(Xi, omega[b], lw); mtv(Xi, vcom, lv);
            const double* bx = m->box[b];
            if (m->viscosity > 0) {
                const double diam = (bx[0] + bx[1] + bx[2]) / 3.0;
                for (int k = 0; k < 3; ++k) {
                    lt[k] += -M_PI * diam * diam * diam * m->viscosity * lw[k];
                    lf[k] += -3.0 * M_PI * diam * m->viscosity * lv[k];
                }
            }
            if (m->density > 0) {
                lf[0] -= 0.5 * m->density * bx[1] * bx[2] * fabs(lv[0]) * lv[0];
                lf[1] -= 0.5 * m->density * bx[0] * bx[2] * fabs(lv[1]) * lv[1];
                lf[2] -= 0.5 * m->density * bx[0] * bx[1] * fabs(lv[2]) * lv[2];
                lt[0] -= m->density * bx[0] * (pow(bx[1], 4) + pow(bx[2], 4)) * fabs(lw[0]) * lw[0] / 64.0;
                lt[1] -= m->density * bx[1] * (pow(bx[0], 4) + pow(bx[2], 4)) * fabs(lw[1]) * lw[1] / 64.0;
                lt[2] -= m->density * bx[2] * (pow(bx[0], 4) + pow(bx[1], 4)) * fabs(lw[2]) * lw[2] / 64.0;
            }
            mv(Xi, lf, fw); mv(Xi, lt, tw);
            for (int v = 0; v < nv; ++v)
                rhs[v] += Jv[0][v] * fw[0] + Jv[1][v] * fw[1] + Jv[2][v] * fw[2] + Jw[0][v] * tw[0] + Jw[1][v] * tw[1] + Jw[2][v] * tw[2];
        }
    }
    for (int k = 0; k < m->nu; ++k) {
        double F = ctrl[k];
        if (m->ctrl_limited[k]) F = fmin(fmax(F, m->ctrl_lo[k]), m->ctrl_hi[k]);
        const int s = m->act_site[k], b = m->site_body[s];
        double t[3], sp[3], sq[4], Rs[9], fw[3], tw[3];
        mv(R[b], m->site_pos[s], t);
        for (int i = 0; i < 3; ++i) sp[i] = xpos[b][i] + t[i];
        qmul(xquat[b], m->site_quat[s], sq);
        q2m(sq, Rs);
        mv(Rs, &m->gear[k][0], fw); mv(Rs, &m->gear[k][3], tw);
        jac(m, &d, b, sp, Jv, Jw);
        for (int v = 0; v < nv; ++v)
            rhs[v] += F * (Jv[0][v] * fw[0] + Jv[1][v] * fw[1] + Jv[2][v] * fw[2] + Jw[0][v] * tw[0] + Jw[1][v] * tw[1] + Jw[2][v] * tw[2]);
    }
    /* implicit Euler damping, then Cholesky solve */
    for (int v = 0; v < nv; ++v) M[v][v] += m->dt * m->damping[v];
    double L[MAXV][MAXV];
    memset(L, 0, sizeof(L));
    for (int i = 0; i < nv; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = M[i][j];
            for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
            L[i][j] = (i == j) ? sqrt(s) : s / L[j][j];
        }
    double y[MAXV], qacc[MAXV];
    for (int i = 0; i < nv; ++i) { double s = rhs[i]; for (int k = 0; k < i; ++k) s -= L[i][k] * y[k]; y[i] = s / L[i][i]; }
    for (int i = nv - 1; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < nv; ++k) s -= L[k][i] * qacc[k]; qacc[i] = s / L[i][i]; }
    for (int v = 0; v < nv; ++v) qvel[v] += m->dt * qacc[v];
    for (int j = 0; j < m->njnt; ++j) {
        const int qa = m->qadr[j], va = m->vadr[j];
        if (m->jtype[j] == JNT_FREE) {
            for (int k = 0; k < 3; ++k) qpos[qa + k] += m->dt * qvel[va + k];
            double q[4] = {qpos[qa + 3], qpos[qa + 4], qpos[qa + 5], qpos[qa + 6]}, dq[4], o[4];
            qnorm(q);
            const double* w = &qvel[va + 3];
            const double n = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
            if (n > 0) {
                const double h = 0.5 * m->dt * n, sn = sin(h) / n;
                dq[0] = cos(h); dq[1] = sn * w[0]; dq[2] = sn * w[1]; dq[3] = sn * w[2];
            } else { dq[0] = 1; dq[1] = dq[2] = dq[3] = 0; }
            qmul(q, dq, o);
            qnorm(o);
            for (int k = 0; k < 4; ++k) qpos[qa + 3 + k] = o[k];
        } else {
            qpos[qa] += m->dt * qvel[va];
        }
    }
}

/* ---- minimal pthread parallel-for (this image's gcc cannot find its OpenMP spec file) --------- */
typedef void (*range_fn)(void* ctx, int lo, int hi, long* acc);
typedef struct { range_fn fn; void* ctx; int lo, hi; long acc; } Task;
static void* task_main(void* p) { Task* t = (Task*)p; t->fn(t->ctx, t->lo, t->hi, &t->acc); return 0; }

int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static long parallel_for(range_fn fn, void* ctx, int n, int threads) {
    if (threads <= 0) threads = orc_max_threads();
    if (threads > n) threads = n > 0 ? n : 1;
    if (threads > 256) threads = 256;
    Task tasks[256];
    pthread_t tid[256];
    long total = 0;
    for (int t = 0; t < threads; ++t) {
        tasks[t].fn = fn; tasks[t].ctx = ctx; tasks[t].acc = 0;
        tasks[t].lo = (int)((long)n * t / threads); tasks[t].hi = (int)((long)n * (t + 1) / threads);
        if (t > 0) pthread_create(&tid[t], 0, task_main, &tasks[t]);
    }
    task_main(&tasks[0]);
    for (int t = 1; t < threads; ++t) pthread_join(tid[t], 0);
    for (int t = 0; t < threads; ++t) total += tasks[t].acc;
    return total;
}

typedef struct { const OrcModel* m; double* qpos; double* qvel; const double* ctrl; } StepCtx;
static void step_range(void* c, int lo, int hi, long* acc) {
    StepCtx* s = (StepCtx*)c;
    (void)acc;
    for (int i = lo; i < hi; ++i)
        orc_step_one(s->m, s->qpos + (size_t)i * s->m->nq, s->qvel + (size_t)i * s->m->nv, s->ctrl + (size_t)i * s->m->nu);
}

void orc_step_batch(const OrcModel* m, int n, double* qpos, double* qvel, const double* ctrl, int threads) {
    StepCtx c = {m, qpos, qvel, ctrl};
    parallel_for(step_range, &c, n, threads);
}

int orc_model_size(void) { return (int)sizeof(OrcModel); }

/* ---- HoverEnv semantics for the timed CPU baseline (envs/hover_env.py:159-238) ----------
 * North-star variant: no battery sag, auto-reset with a cheap LCG (the CPU baseline is timed,
 * not parity-checked on its reset stream; parity uses oracle/envs.py + oracle/philox.py). */
typedef struct {
    double act_lo[4], act_hi[4], mix_inv[16], max_thrust;
    double obs_lo[12], obs_hi[12], term_lo[12], term_hi[12], init_lo[12], init_hi[12], tgt_lo[3], tgt_hi[3];
    int max_episode_steps;
} OrcHover;

static double lcg01(uint64_t* s) {
    *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (double)(*s >> 11) * (1.0 / 9007199254740992.0);
}

static void quat_to_rpy(const double q[4], double rpy[3]) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    double sp = 2 * (w * y - x * z);
    sp = sp > 1 ? 1 : (sp < -1 ? -1 : sp);
    rpy[0] = atan2(2 * (y * z + w * x), 1 - 2 * (x * x + y * y));
    rpy[1] = asin(sp);
    rpy[2] = atan2(2 * (x * y + w * z), 1 - 2 * (y * y + z * z));
}

static void hover_reset(const OrcHover* h, uint64_t* rng, double* qpos, double* qvel, double* tgt, int* count) {
    double s[12];
    for (int i = 0; i < 12; ++i) s[i] = h->init_lo[i] + lcg01(rng) * (h->init_hi[i] - h->init_lo[i]);
    for (int i = 0; i < 3; ++i) tgt[i] = h->tgt_lo[i] + lcg01(rng) * (h->tgt_hi[i] - h->tgt_lo[i]);
    const double cr = cos(0.5 * s[3]), sr = sin(0.5 * s[3]), cp = cos(0.5 * s[4]), sp = sin(0.5 * s[4]);
    const double cy = cos(0.5 * s[5]), sy = sin(0.5 * s[5]);
    memset(qpos, 0, 11 * sizeof(double)); memset(qvel, 0, 10 * sizeof(double));
    qpos[0] = s[0]; qpos[1] = s[1]; qpos[2] = s[2];
    qpos[3] = cr * cp * cy + sr * sp * sy; qpos[4] = sr * cp * cy - cr * sp * sy;
    qpos[5] = cr * sp * cy + sr * cp * sy; qpos[6] = cr * cp * sy - sr * sp * cy;
    for (int i = 0; i < 6; ++i) qvel[i] = s[6 + i];
    *count = 0;
}

/* Runs `steps` env steps on n envs with uniformly random actions; returns the number of episodes
 * finished (so the work cannot be optimised away) and writes the last obs/reward. */
typedef struct {
    const OrcModel* m; const OrcHover* h; int steps; uint64_t seed; double* qpos; double* qvel; double* tgt; int* count;
    float* obs; float* reward; int do_reset;
} HoverCtx;

static void hover_range(void* c, int lo, int hi, long* acc) {
    HoverCtx* x = (HoverCtx*)c;
    const OrcModel* m = x->m; const OrcHover* h = x->h;
    long finished = 0;
    for (int i = lo; i < hi; ++i) {
        uint64_t rng = x->seed * 0x9E3779B97F4A7C15ULL + (uint64_t)i * 0xD1B54A32D192ED03ULL + 1;
        double* qp = x->qpos + (size_t)i * 11; double* qv = x->qvel + (size_t)i * 10; double* tg = x->tgt + (size_t)i * 3;
        int* count = x->count;
        float* obs = x->obs; float* reward = x->reward;
        if (x->do_reset) hover_reset(h, &rng, qp, qv, tg, &count[i]);
        for (int t = 0; t < x->steps; ++t) {
            double u[4], ctrl[4];
            for (int k = 0; k < 4; ++k) {
                const double a = 2.0 * lcg01(&rng) - 1.0;
                u[k] = (a + 1.0) / 2.0 * (h->act_hi[k] - h->act_lo[k]) + h->act_lo[k];
            }
            for (int k = 0; k < 4; ++k) {
                double F = h->mix_inv[4 * k] * u[0] + h->mix_inv[4 * k + 1] * u[1] + h->mix_inv[4 * k + 2] * u[2] + h->mix_inv[4 * k + 3] * u[3];
                ctrl[k] = fmin(fmax(F, 0.0), h->max_thrust);
            }
            orc_step_one(m, qp, qv, ctrl);
            count[i] += 1;
            double rpy[3];
            quat_to_rpy(qp + 3, rpy);
            float s12[12] = {(float)qp[0], (float)qp[1], (float)qp[2], (float)rpy[0], (float)rpy[1], (float)rpy[2],
                             (float)qv[0], (float)qv[1], (float)qv[2], (float)qv[3], (float)qv[4], (float)qv[5]};
            float xx[12];
            for (int k = 0; k < 12; ++k) xx[k] = s12[k];
            for (int k = 0; k < 3; ++k) xx[k] = (float)tg[k] - s12[k];
            int inside = 1;
            for (int k = 0; k < 12; ++k) {
                obs[(size_t)i * 12 + k] = 2.0f * (xx[k] - (float)h->obs_lo[k]) / ((float)h->obs_hi[k] - (float)h->obs_lo[k]) - 1.0f;
                if (!(isfinite(s12[k]) && s12[k] >= (float)h->term_lo[k] && s12[k] <= (float)h->term_hi[k])) inside = 0;
            }
            const double dx = s12[0] - (float)tg[0], dy = s12[1] - (float)tg[1], dz = s12[2] - (float)tg[2];
            reward[i] = (float)exp(-(dx * dx + dy * dy + dz * dz));
            if (!inside || count[i] >= h->max_episode_steps) {
                finished += 1;
                hover_reset(h, &rng, qp, qv, tg, &count[i]);
            }
        }
    }
    *acc = finished;
}

/* Runs `steps` env steps on n envs with uniformly random actions; returns the number of episodes
 * finished (so the work cannot be optimised away) and writes the last obs/reward. */
long orc_hover_rollout(const OrcModel* m, const OrcHover* h, int n, int steps, uint64_t seed, double* qpos, double* qvel,
                       double* tgt, int* count, float* obs, float* reward, int threads, int do_reset) {
    HoverCtx c = {m, h, steps, seed, qpos, qvel, tgt, count, obs, reward, do_reset};
    return parallel_for(hover_range, &c, n, threads);
}
