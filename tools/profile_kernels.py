"""One launch of every hot kernel at (reduced-T) bench sizes, for `ncu --set full -k regex:<name>` captures.

    ncu --set full --clock-control none --import-source on -k regex:rollout_policy_tc -c 2 -o out python tools/profile_kernels.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import make_policy_params
from uav_reinforcement_learning_control_b200 import config as Q
from uav_reinforcement_learning_control_b200.engine import Engine

dev = torch.device("cuda", 0)
which = set(sys.argv[1:]) or {"step", "resident", "fma", "tc", "tc_large", "gae", "tc21"}

if {"step", "resident"} & which:
    n = 1 << 20
    eng = Engine(Q.EnvConfig.north_star(seed=0), n, device=0)
    st = eng.new_state(); eng.reset(st)
    act = torch.rand(n, 4, device=dev) * 2 - 1
    if "step" in which:
        for _ in range(4):
            eng.step(st, act)
    if "resident" in which:
        stats = torch.zeros(4, n, device=dev)
        eng.rollout_random(st, 64, t0=0, stats=stats)
    torch.cuda.synchronize(); del eng

if {"fma", "tc", "gae"} & which:
    nb, T = 8192, 256
    eng = Engine(Q.EnvConfig.north_star(seed=1), nb, device=0)
    st = eng.new_state(); eng.reset(st)
    params = make_policy_params(eng, torch, dev, seed=0)
    buf = None
    if "fma" in which:
        buf = eng.rollout_policy(st, params, T=T, t0=0, dist=0)
    if "tc" in which or buf is None:
        buf = eng.rollout_policy(st, params, T=T, t0=T, dist=0, buffers=buf, tensor_cores=True)
        buf = eng.rollout_policy(st, params, T=T, t0=2 * T, dist=0, buffers=buf, tensor_cores=True)
    if "gae" in which:
        big = {k: buf[k].repeat(4, 1) for k in ("reward", "value", "done", "trunc")}        # 1024 x 8192
        eng.gae(big["reward"], big["value"], big["done"], big["trunc"], buf["last_value"], 0.99, 0.95)
    torch.cuda.synchronize(); del eng

if "tc_large" in which:
    nb, T = 1 << 18, 32
    eng = Engine(Q.EnvConfig.north_star(seed=2), nb, device=0)
    st = eng.new_state(); eng.reset(st)
    params = make_policy_params(eng, torch, dev, seed=0)
    buf = eng.rollout_policy(st, params, T=T, t0=0, dist=0, tensor_cores=True)
    buf = eng.rollout_policy(st, params, T=T, t0=T, dist=0, buffers=buf, tensor_cores=True)
    torch.cuda.synchronize(); del eng

if "tc21" in which:
    nb, T = 8192, 128
    eng = Engine(Q.EnvConfig.mjx_brax(episode_length=500, auto_reset=Q.RESET_RESTORE_FIRST, seed=3), nb, device=0)
    st = eng.new_state(); first = torch.zeros(21, nb, device=dev); eng.reset(st, first_state=first)
    params = make_policy_params(eng, torch, dev, seed=0, dist=1)
    buf = eng.rollout_policy(st, params, T=T, t0=0, dist=1, first_state=first, tensor_cores=True)
    torch.cuda.synchronize(); del eng
print("profile_kernels done")
