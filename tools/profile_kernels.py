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
which = set(sys.argv[1:]) or {"step", "resident", "fma", "tc", "tc_large", "gae", "tc21", "ppo", "ppo21"}

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
if "ppo" in which:
    # two minibatch updates of the configs[2] shape: 2^20 random rows out of 8192 x 1024 samples
    from uav_reinforcement_learning_control_b200.ppo import ActorCritic, FusedUpdater
    N = 8192 * 1024; mb = N // 8
    g = torch.Generator(device=dev); g.manual_seed(0)
    obs = torch.rand(N, 12, device=dev, generator=g) * 2 - 1
    act = torch.randn(N, 4, device=dev, generator=g) * 0.4
    old_logp = torch.randn(N, device=dev, generator=g) * 0.1 - 1.0
    adv = torch.randn(N, device=dev, generator=g); ret = torch.randn(N, device=dev, generator=g)
    params = ActorCritic(12, dev, seed=0, log_std_init=-1.0).pack()
    up = FusedUpdater(dev)
    perm = torch.randperm(N, device=dev, generator=g).to(torch.int32)
    packed = up.pack(obs, act, old_logp, adv, ret)
    for k in range(2):
        up.grad(params, adv=adv, packed=packed, idx=perm[k * mb:(k + 1) * mb], clip_range=0.19, vf_coef=0.5, ent_coef=1e-4)
        up.adam(params, 1.5e-4)
    torch.cuda.synchronize()
    del obs, act, old_logp, adv, ret, packed, perm
if "ppo21" in which:
    # the Brax policy's update (21-D obs, tanh-normal head) through the single-tile kernel: two 2^18-row minibatches
    from uav_reinforcement_learning_control_b200.ppo import FusedUpdater, init_packed_params
    N = 8192 * 256; mb = N // 8
    g = torch.Generator(device=dev); g.manual_seed(0)
    obs = torch.rand(N, 21, device=dev, generator=g) * 2 - 1
    act = torch.randn(N, 4, device=dev, generator=g) * 0.4
    old_logp = torch.randn(N, device=dev, generator=g) * 0.1 - 1.0
    adv = torch.randn(N, device=dev, generator=g); ret = torch.randn(N, device=dev, generator=g)
    params = init_packed_params(21, 1, 0).to(dev)
    up = FusedUpdater(dev, obs_dim=21, dist=1)
    perm = torch.randperm(N, device=dev, generator=g).to(torch.int32)
    packed = up.pack(obs, act, old_logp, adv, ret)
    for k in range(2):
        up.grad(params, adv=adv, packed=packed, idx=perm[k * mb:(k + 1) * mb], clip_range=0.3, vf_coef=0.25, ent_coef=1e-3,
                normalize_adv=2, sample_seed=k)
        up.adam(params, 3e-4, max_grad_norm=0.0, eps=1e-8)
    torch.cuda.synchronize()
print("profile_kernels done")
