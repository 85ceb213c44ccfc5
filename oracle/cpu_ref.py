"""ORACLE (test infrastructure, not product code): ctypes binding of oracle/cpu_ref.c.

Used (a) by tests to cross-check the C restatement against the NumPy one and (b) by bench.py's
cpu_baseline / `--impl reference` arm as the timed CPU port ("kind": "port").
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MAXB, MAXV, MAXU = 8, 16, 8


class OrcModel(C.Structure):
    _fields_ = [
        ("nb", C.c_int), ("njnt", C.c_int), ("nq", C.c_int), ("nv", C.c_int), ("nu", C.c_int), ("nsite", C.c_int),
        ("parent", C.c_int * MAXB),
        ("pos", (C.c_double * 3) * MAXB), ("quat", (C.c_double * 4) * MAXB), ("ipos", (C.c_double * 3) * MAXB),
        ("iquat", (C.c_double * 4) * MAXB), ("mass", C.c_double * MAXB), ("inertia", (C.c_double * 3) * MAXB),
        ("jtype", C.c_int * MAXB), ("jbody", C.c_int * MAXB), ("qadr", C.c_int * MAXB), ("vadr", C.c_int * MAXB),
        ("jpos", (C.c_double * 3) * MAXB), ("jaxis", (C.c_double * 3) * MAXB),
        ("damping", C.c_double * MAXV), ("armature", C.c_double * MAXV),
        ("site_body", C.c_int * MAXB), ("site_pos", (C.c_double * 3) * MAXB), ("site_quat", (C.c_double * 4) * MAXB),
        ("act_site", C.c_int * MAXU), ("ctrl_limited", C.c_int * MAXU),
        ("gear", (C.c_double * 6) * MAXU), ("ctrl_lo", C.c_double * MAXU), ("ctrl_hi", C.c_double * MAXU),
        ("dt", C.c_double), ("gravity", C.c_double * 3), ("density", C.c_double), ("viscosity", C.c_double),
        ("box", (C.c_double * 3) * MAXB),
    ]


class OrcHover(C.Structure):
    _fields_ = [("act_lo", C.c_double * 4), ("act_hi", C.c_double * 4), ("mix_inv", C.c_double * 16),
                ("max_thrust", C.c_double), ("obs_lo", C.c_double * 12), ("obs_hi", C.c_double * 12),
                ("term_lo", C.c_double * 12), ("term_hi", C.c_double * 12), ("init_lo", C.c_double * 12),
                ("init_hi", C.c_double * 12), ("tgt_lo", C.c_double * 3), ("tgt_hi", C.c_double * 3),
                ("max_episode_steps", C.c_int)]


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "libcpu_ref.so")
        src = os.path.join(HERE, "cpu_ref.c")
        if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
            subprocess.check_call(["make", "-s", "-C", HERE])
        _LIB = C.CDLL(so)
        assert _LIB.orc_model_size() == C.sizeof(OrcModel)
        _LIB.orc_hover_rollout.restype = C.c_long
    return _LIB


def pack_model(tree) -> OrcModel:
    from .mujoco_pipeline import inertia_box, JNT_FREE
    m = OrcModel()
    m.nb, m.njnt, m.nq, m.nv, m.nu, m.nsite = tree.nbody, len(tree.jnt_type), tree.nq, tree.nv, tree.nu, len(tree.site_body)
    assert m.nb <= MAXB and m.nv <= MAXV and m.nu <= MAXU and m.nsite <= MAXB
    q = v = 0
    for b in range(tree.nbody):
        m.parent[b] = tree.body_parent[b]; m.mass[b] = tree.body_mass[b]
        for k in range(3):
            m.pos[b][k] = tree.body_pos[b][k]; m.ipos[b][k] = tree.body_ipos[b][k]; m.inertia[b][k] = tree.body_inertia[b][k]
        for k in range(4):
            m.quat[b][k] = tree.body_quat[b][k]; m.iquat[b][k] = tree.body_iquat[b][k]
        if tree.body_mass[b] > 0:
            bx = inertia_box(tree.body_mass[b], tree.body_inertia[b])
            for k in range(3):
                m.box[b][k] = bx[k]
    for j, ty in enumerate(tree.jnt_type):
        m.jtype[j], m.jbody[j], m.qadr[j], m.vadr[j] = ty, tree.jnt_body[j], q, v
        nvj = 6 if ty == JNT_FREE else 1
        for k in range(3):
            m.jpos[j][k] = tree.jnt_pos[j][k]; m.jaxis[j][k] = tree.jnt_axis[j][k]
        for k in range(nvj):
            m.damping[v + k] = tree.jnt_damping[j]; m.armature[v + k] = tree.jnt_armature[j]
        q += 7 if ty == JNT_FREE else 1
        v += nvj
    for s in range(m.nsite):
        m.site_body[s] = tree.site_body[s]
        for k in range(3):
            m.site_pos[s][k] = tree.site_pos[s][k]
        for k in range(4):
            m.site_quat[s][k] = tree.site_quat[s][k]
    for a in range(tree.nu):
        m.act_site[a] = tree.act_site[a]
        for k in range(6):
            m.gear[a][k] = tree.act_gear[a][k]
        if tree.act_ctrlrange[a] is not None:
            m.ctrl_limited[a] = 1; m.ctrl_lo[a], m.ctrl_hi[a] = tree.act_ctrlrange[a]
    m.dt = tree.timestep; m.density = tree.density; m.viscosity = tree.viscosity
    for k in range(3):
        m.gravity[k] = tree.gravity[k]
    return m


def pack_hover(cfg) -> OrcHover:
    h = OrcHover()
    for name, src in (("act_lo", cfg.act_lo), ("act_hi", cfg.act_hi), ("mix_inv", cfg.mixer()[1].ravel()),
                      ("obs_lo", np.float32(cfg.obs_lo)), ("obs_hi", np.float32(cfg.obs_hi)),
                      ("term_lo", np.float32(cfg.term_lo)), ("term_hi", np.float32(cfg.term_hi)),
                      ("init_lo", cfg.init_lo), ("init_hi", cfg.init_hi), ("tgt_lo", cfg.target_lo), ("tgt_hi", cfg.target_hi)):
        arr = getattr(h, name)
        for i, v in enumerate(np.asarray(src, dtype=np.float64)):
            arr[i] = float(v)
    h.max_thrust = cfg.max_motor_thrust
    h.max_episode_steps = cfg.max_episode_steps
    return h


def step_batch(tree, qpos, qvel, ctrl, threads=1):
    """C oracle: one physics step on [n, nq] / [n, nv] / [n, nu] float64 arrays (returns new copies)."""
    m = pack_model(tree)
    qpos = np.ascontiguousarray(qpos, dtype=np.float64).copy(); qvel = np.ascontiguousarray(qvel, dtype=np.float64).copy()
    ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib().orc_step_batch(C.byref(m), qpos.shape[0], p(qpos), p(qvel), p(ctrl), int(threads))
    return qpos, qvel


class HoverRollout:
    """Timed CPU port of the hover workload (random actions, auto-reset), all host threads."""

    def __init__(self, tree, cfg, n):
        self.m = pack_model(tree); self.h = pack_hover(cfg); self.n = n
        self.qpos = np.zeros((n, 11)); self.qvel = np.zeros((n, 10)); self.tgt = np.zeros((n, 3))
        self.count = np.zeros(n, np.int32); self.obs = np.zeros((n, 12), np.float32); self.rew = np.zeros(n, np.float32)
        self.first = True

    def run(self, steps, seed=0, threads=0):
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        fin = lib().orc_hover_rollout(C.byref(self.m), C.byref(self.h), self.n, int(steps), C.c_uint64(seed), p(self.qpos),
                                      p(self.qvel), p(self.tgt), p(self.count), p(self.obs), p(self.rew), int(threads),
                                      1 if self.first else 0)
        self.first = False
        return int(fin)

    @staticmethod
    def max_threads():
        return int(lib().orc_max_threads())
