"""ORACLE (test infrastructure, not product code) -- PARITY UNPINNED.

A float64 NumPy restatement of what ``mjx.step`` / ``mujoco.mj_step`` compute for
the reference's model (reference call sites: train_brax_ppo.py:317,
envs/jax_mjx_quad_env.py:144, envs/hover_env.py:180,
envs/trajectory_follow_env.py:154; model: model/drone/drone.xml).

The arithmetic lives in third-party MuJoCo / MJX, which is neither vendored under
/root/reference, nor version-pinned by it, nor installable here.  This file
restates MuJoCo's *published* forward-dynamics pipeline for a generic kinematic
tree of free + hinge joints:

    kinematics -> per-body Jacobians -> joint-space inertia M (what CRB yields)
    -> bias forces (what RNE yields, gravity included) -> passive forces
    (joint damping; inertia-box fluid model, engine_passive.c
    ``mj_inertiaBoxFluidModel`` / mjx passive.py ``_inertia_box_fluid_model``)
    -> motor forces through site transmissions (ctrl clamped to ctrlrange)
    -> qacc = M^-1 (passive + actuator - bias)
    -> semi-implicit Euler with implicit joint damping (``mj_EulerSkip`` / mjx
    ``euler``): qvel += dt*qacc, then qpos integrated with the NEW qvel; free-joint
    quaternion advanced by the body-frame angular velocity (``mju_quatIntegrate``
    / mjx ``math.quat_integrate``) and renormalised.

It is driven purely by the parsed tree (``TreeModel``): it knows nothing about
quadrotors.  The CUDA kernels integrate a hand-derived closed form of the same
system, so agreement between the two is an independent check of both.

"Parity unpinned": the reference holds no golden vectors for this path (its three
test_*.py scripts only print) and MuJoCo cannot be executed here, so nothing pins
this restatement to MuJoCo's actual output.  What *is* pinned: analytic known
answers, conservation laws, and tests/golden vectors generated from the
reference's importable utils (see tests/golden/make_golden.py).

Conventions (MuJoCo): free joint qpos = [pos(3, world), quat(4, wxyz)], qvel =
[linear velocity of the body origin (world axes), angular velocity (BODY axes)].
"""
from __future__ import annotations

import numpy as np

JNT_FREE, JNT_HINGE = 0, 3


# ---------------------------------------------------------------- quaternions
def quat_mul(a, b):
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ], axis=-1)


def quat_to_mat(q):
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3), dtype=q.dtype)
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - w * z); R[..., 0, 2] = 2 * (x * z + w * y)
    R[..., 1, 0] = 2 * (x * y + w * z); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - w * x)
    R[..., 2, 0] = 2 * (x * z - w * y); R[..., 2, 1] = 2 * (y * z + w * x); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def axis_angle_quat(axis, angle):
    h = 0.5 * angle
    return np.concatenate([np.cos(h)[..., None], np.sin(h)[..., None] * axis], axis=-1)


def normalize(v):
    n = np.linalg.norm(v, axis=-1, keepdims=True)
    return v / n


def inertia_box(mass, inertia):
    I0, I1, I2 = inertia
    return np.array([
        np.sqrt(max(1e-15, I1 + I2 - I0) / mass * 6.0),
        np.sqrt(max(1e-15, I0 + I2 - I1) / mass * 6.0),
        np.sqrt(max(1e-15, I0 + I1 - I2) / mass * 6.0),
    ])


# ---------------------------------------------------------------- the pipeline
class TreePipeline:
    """Forward dynamics + Euler step for a TreeModel, batched over the leading axis."""

    def __init__(self, tree, fluid=True, gravity=True):
        self.t = tree
        self.nq, self.nv, self.nu, self.nb = tree.nq, tree.nv, tree.nu, tree.nbody
        self.dt = float(tree.timestep)
        self.gravity = np.asarray(tree.gravity, dtype=np.float64) * (1.0 if gravity else 0.0)
        self.fluid = fluid and (tree.density > 0 or tree.viscosity > 0)
        # address maps
        self.qadr, self.vadr = [], []
        q = v = 0
        for ty in tree.jnt_type:
            self.qadr.append(q); self.vadr.append(v)
            q += 7 if ty == JNT_FREE else 1
            v += 6 if ty == JNT_FREE else 1
        self.body_jnt = {b: [j for j in range(len(tree.jnt_type)) if tree.jnt_body[j] == b] for b in range(self.nb)}
        damp = np.zeros(self.nv); arm = np.zeros(self.nv)
        for j, ty in enumerate(tree.jnt_type):
            n = 6 if ty == JNT_FREE else 1
            damp[self.vadr[j]:self.vadr[j] + n] = tree.jnt_damping[j]
            arm[self.vadr[j]:self.vadr[j] + n] = tree.jnt_armature[j]
        self.damping, self.armature = damp, arm
        self.boxes = [None if tree.body_mass[b] <= 0 else inertia_box(tree.body_mass[b], tree.body_inertia[b])
                      for b in range(self.nb)]

    # -- kinematics + Jacobians ------------------------------------------------
    def kinematics(self, qpos):
        """Returns per-body world frames and the Jacobian columns' (axis, anchor, kind) lists."""
        t = self.t
        B = qpos.shape[0]
        xpos = [np.zeros((B, 3))]; xquat = [np.tile(np.array([1.0, 0, 0, 0]), (B, 1))]
        dofs = [[]]      # per body: list of (vadr, kind, axis_world[B,3], anchor_world[B,3])
        for b in range(1, self.nb):
            p = t.body_parent[b]
            Rp = quat_to_mat(xquat[p])
            pos = xpos[p] + np.einsum("bij,j->bi", Rp, t.body_pos[b])
            quat = quat_mul(xquat[p], np.broadcast_to(t.body_quat[b], (B, 4)))
            mydofs = list(dofs[p])
            for j in self.body_jnt[b]:
                qa, va = self.qadr[j], self.vadr[j]
                if t.jnt_type[j] == JNT_FREE:
                    pos = qpos[:, qa:qa + 3].copy()
                    quat = normalize(qpos[:, qa + 3:qa + 7])      # used normalised, not written back
                    R = quat_to_mat(quat)
                    for k in range(3):
                        ax = np.zeros((B, 3)); ax[:, k] = 1.0
                        mydofs.append((va + k, "lin", ax, None))
                    for k in range(3):
                        mydofs.append((va + 3 + k, "rot", R[:, :, k].copy(), pos))
                else:
                    R = quat_to_mat(quat)
                    anchor = pos + np.einsum("bij,j->bi", R, t.jnt_pos[j])
                    axis_l = t.jnt_axis[j] / np.linalg.norm(t.jnt_axis[j])
                    axis_w = np.einsum("bij,j->bi", R, axis_l)
                    quat = quat_mul(quat, axis_angle_quat(np.broadcast_to(axis_l, (B, 3)), qpos[:, qa]))
                    R2 = quat_to_mat(quat)
                    pos = anchor - np.einsum("bij,j->bi", R2, t.jnt_pos[j])
                    mydofs.append((va, "rot", axis_w, anchor))
            xpos.append(pos); xquat.append(quat); dofs.append(mydofs)
        return xpos, xquat, dofs

    def jacobians(self, point, dofs_b, B):
        """Linear Jacobian of a world point rigidly attached to the body, and its angular Jacobian."""
        Jv = np.zeros((B, 3, self.nv)); Jw = np.zeros((B, 3, self.nv))
        for va, kind, ax, anchor in dofs_b:
            if kind == "lin":
                Jv[:, :, va] = ax
            else:
                Jw[:, :, va] = ax
                Jv[:, :, va] = np.cross(ax, point - anchor)
        return Jv, Jw

    # -- forward dynamics ------------------------------------------------------
    def forward(self, qpos, qvel, ctrl, return_parts=False):
        t = self.t
        qpos = np.asarray(qpos, dtype=np.float64); qvel = np.asarray(qvel, dtype=np.float64)
        ctrl = np.asarray(ctrl, dtype=np.float64)
        B = qpos.shape[0]
        xpos, xquat, dofs = self.kinematics(qpos)
        M = np.zeros((B, self.nv, self.nv)); bias = np.zeros((B, self.nv)); passive = np.zeros((B, self.nv))
        M[:, np.arange(self.nv), np.arange(self.nv)] += self.armature
        passive -= self.damping * qvel

        # bias needs velocity-product accelerations: recurse body by body (world frame)
        omega = [np.zeros((B, 3))]; alpha0 = [np.zeros((B, 3))]
        vorg = [np.zeros((B, 3))]; aorg0 = [np.zeros((B, 3))]      # body-origin velocity / vel-product accel
        for b in range(1, self.nb):
            p = t.body_parent[b]
            Rp = quat_to_mat(xquat[p])
            d = xpos[b] - xpos[p]
            # start from the parent's motion carried to this body's origin
            w = omega[p].copy(); al = alpha0[p].copy()
            v = vorg[p] + np.cross(omega[p], d)
            a = aorg0[p] + np.cross(alpha0[p], d) + np.cross(omega[p], np.cross(omega[p], d))
            for j in self.body_jnt[b]:
                va = self.vadr[j]
                if t.jnt_type[j] == JNT_FREE:
                    R = quat_to_mat(xquat[b])
                    # world-axis linear dofs and body-axis angular dofs: with qacc = 0 the origin does not
                    # accelerate and d/dt(R w_body) = (R w_body) x (R w_body) = 0
                    v = qvel[:, va:va + 3].copy()
                    w = np.einsum("bij,bj->bi", R, qvel[:, va + 3:va + 6])
                    a = np.zeros((B, 3)); al = np.zeros((B, 3))
                else:
                    # hinge: find axis/anchor recorded by kinematics
                    ax, anchor = [(x[2], x[3]) for x in dofs[b] if x[0] == va][0]
                    s = qvel[:, va][:, None]
                    r = xpos[b] - anchor                      # origin relative to the anchor
                    # motion of the anchor (attached to the parent side)
                    da = anchor - xpos[p]
                    v_anchor = vorg[p] + np.cross(omega[p], da)
                    a_anchor = aorg0[p] + np.cross(alpha0[p], da) + np.cross(omega[p], np.cross(omega[p], da))
                    w_new = w + s * ax
                    al = al + np.cross(w, s * ax)             # d/dt(s*axis) with ds/dt = 0
                    v = v_anchor + np.cross(w_new, r)
                    a = a_anchor + np.cross(al, r) + np.cross(w_new, np.cross(w_new, r))
                    w = w_new
            omega.append(w); alpha0.append(al); vorg.append(v); aorg0.append(a)

        parts = {}
        for b in range(1, self.nb):
            mb = t.body_mass[b]
            if mb <= 0:
                continue
            R = quat_to_mat(xquat[b])
            xipos = xpos[b] + np.einsum("bij,j->bi", R, t.body_ipos[b])
            ximat = quat_to_mat(quat_mul(xquat[b], np.broadcast_to(t.body_iquat[b], (B, 4))))
            Iw = np.einsum("bij,j,bkj->bik", ximat, np.asarray(t.body_inertia[b], dtype=np.float64), ximat)
            Jv, Jw = self.jacobians(xipos, dofs[b], B)
            M += mb * np.einsum("bki,bkj->bij", Jv, Jv) + np.einsum("bki,bkl,blj->bij", Jw, Iw, Jw)
            rc = xipos - xpos[b]
            a_com = aorg0[b] + np.cross(alpha0[b], rc) + np.cross(omega[b], np.cross(omega[b], rc))
            f = mb * (a_com - self.gravity)
            n = np.einsum("bij,bj->bi", Iw, alpha0[b]) + np.cross(omega[b], np.einsum("bij,bj->bi", Iw, omega[b]))
            bias += np.einsum("bki,bk->bi", Jv, f) + np.einsum("bki,bk->bi", Jw, n)
            if self.fluid:
                v_com = vorg[b] + np.cross(omega[b], rc)
                lw = np.einsum("bji,bj->bi", ximat, omega[b])         # local angular velocity
                lv = np.einsum("bji,bj->bi", ximat, v_com)            # local linear velocity at the COM
                box = self.boxes[b]
                lt = np.zeros((B, 3)); lf = np.zeros((B, 3))
                if t.viscosity > 0:
                    diam = float(np.mean(box))
                    lt += -np.pi * diam ** 3 * t.viscosity * lw
                    lf += -3.0 * np.pi * diam * t.viscosity * lv
                if t.density > 0:
                    b0, b1, b2 = box
                    lf[:, 0] -= 0.5 * t.density * b1 * b2 * np.abs(lv[:, 0]) * lv[:, 0]
                    lf[:, 1] -= 0.5 * t.density * b0 * b2 * np.abs(lv[:, 1]) * lv[:, 1]
                    lf[:, 2] -= 0.5 * t.density * b0 * b1 * np.abs(lv[:, 2]) * lv[:, 2]
                    lt[:, 0] -= t.density * b0 * (b1 ** 4 + b2 ** 4) * np.abs(lw[:, 0]) * lw[:, 0] / 64.0
                    lt[:, 1] -= t.density * b1 * (b0 ** 4 + b2 ** 4) * np.abs(lw[:, 1]) * lw[:, 1] / 64.0
                    lt[:, 2] -= t.density * b2 * (b0 ** 4 + b1 ** 4) * np.abs(lw[:, 2]) * lw[:, 2] / 64.0
                fw = np.einsum("bij,bj->bi", ximat, lf); tw = np.einsum("bij,bj->bi", ximat, lt)
                passive += np.einsum("bki,bk->bi", Jv, fw) + np.einsum("bki,bk->bi", Jw, tw)

        # motors through site transmissions
        act = np.zeros((B, self.nv))
        force = ctrl.copy()
        for k in range(self.nu):
            if t.act_ctrlrange[k] is not None:
                force[:, k] = np.clip(ctrl[:, k], t.act_ctrlrange[k][0], t.act_ctrlrange[k][1])
            s = t.act_site[k]; b = t.site_body[s]
            R = quat_to_mat(xquat[b])
            spos = xpos[b] + np.einsum("bij,j->bi", R, t.site_pos[s])
            Rs = quat_to_mat(quat_mul(xquat[b], np.broadcast_to(t.site_quat[s], (B, 4))))
            Jv, Jw = self.jacobians(spos, dofs[b], B)
            fw = np.einsum("bij,j->bi", Rs, t.act_gear[k][:3]) * force[:, k:k + 1]
            tw = np.einsum("bij,j->bi", Rs, t.act_gear[k][3:]) * force[:, k:k + 1]
            act += np.einsum("bki,bk->bi", Jv, fw) + np.einsum("bki,bk->bi", Jw, tw)

        qfrc = passive + act - bias
        if return_parts:
            parts.update(M=M, bias=bias, passive=passive, actuator=act, xpos=xpos, xquat=xquat,
                         omega=omega, vorg=vorg)
            return qfrc, M, parts
        return qfrc, M

    def step(self, qpos, qvel, ctrl):
        """One mj_step / mjx.step: returns (qpos', qvel')."""
        qpos = np.asarray(qpos, dtype=np.float64); qvel = np.asarray(qvel, dtype=np.float64)
        qfrc, M = self.forward(qpos, qvel, ctrl)
        if np.any(self.damping > 0):
            M = M.copy()
            M[:, np.arange(self.nv), np.arange(self.nv)] += self.dt * self.damping
        qacc = np.linalg.solve(M, qfrc[..., None])[..., 0]
        qvel2 = qvel + self.dt * qacc
        return self.integrate_pos(qpos, qvel2), qvel2

    def integrate_pos(self, qpos, qvel):
        t = self.t
        out = qpos.copy()
        for j, ty in enumerate(t.jnt_type):
            qa, va = self.qadr[j], self.vadr[j]
            if ty == JNT_FREE:
                out[:, qa:qa + 3] = qpos[:, qa:qa + 3] + self.dt * qvel[:, va:va + 3]
                w = qvel[:, va + 3:va + 6]
                n = np.linalg.norm(w, axis=-1, keepdims=True)
                axis = np.where(n > 0, w / np.where(n > 0, n, 1.0), 0.0)
                dq = axis_angle_quat(axis, self.dt * n[:, 0])
                out[:, qa + 3:qa + 7] = normalize(quat_mul(normalize(qpos[:, qa + 3:qa + 7]), dq))
            else:
                out[:, qa] = qpos[:, qa] + self.dt * qvel[:, va]
        return out

    # -- diagnostics used by the oracle self-checks ---------------------------
    def momenta(self, qpos, qvel):
        """Total kinetic energy, linear momentum and angular momentum about the world origin."""
        t = self.t
        qpos = np.asarray(qpos, dtype=np.float64); qvel = np.asarray(qvel, dtype=np.float64)
        B = qpos.shape[0]
        xpos, xquat, dofs = self.kinematics(qpos)
        KE = np.zeros(B); P = np.zeros((B, 3)); L = np.zeros((B, 3))
        for b in range(1, self.nb):
            mb = t.body_mass[b]
            if mb <= 0:
                continue
            R = quat_to_mat(xquat[b])
            xipos = xpos[b] + np.einsum("bij,j->bi", R, t.body_ipos[b])
            ximat = quat_to_mat(quat_mul(xquat[b], np.broadcast_to(t.body_iquat[b], (B, 4))))
            Iw = np.einsum("bij,j,bkj->bik", ximat, np.asarray(t.body_inertia[b], dtype=np.float64), ximat)
            Jv, Jw = self.jacobians(xipos, dofs[b], B)
            v = np.einsum("bij,bj->bi", Jv, qvel); w = np.einsum("bij,bj->bi", Jw, qvel)
            Iwv = np.einsum("bij,bj->bi", Iw, w)
            KE += 0.5 * mb * np.sum(v * v, -1) + 0.5 * np.sum(w * Iwv, -1)
            P += mb * v
            L += Iwv + mb * np.cross(xipos, v)
        return KE, P, L
