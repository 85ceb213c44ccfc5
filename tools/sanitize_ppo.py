"""Small PPO-update run (qs_ppo_permutation / qs_ppo_grad / qs_ppo_adam, both schedules) for compute-sanitizer."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from uav_reinforcement_learning_control_b200.ppo import ActorCritic, FusedUpdater

dev = "cuda:0"
g = torch.Generator(device=dev); g.manual_seed(0)
for N, n_idx in ((300, None), (1000, 777), (40000, None)):
    obs = torch.rand(N, 12, device=dev, generator=g) * 2 - 1
    act = torch.randn(N, 4, device=dev, generator=g) * 0.4
    old_logp = torch.randn(N, device=dev, generator=g) * 0.1 - 1.0
    adv = torch.randn(N, device=dev, generator=g); ret = torch.randn(N, device=dev, generator=g)
    params = ActorCritic(12, dev, seed=0, log_std_init=-1.0).pack()
    up = FusedUpdater(dev)
    idx = None if n_idx is None else up.permutation(N, 5, 1)[:n_idx].contiguous()
    for _ in range(2):
        up.grad(params, obs, act, old_logp, adv, ret, idx=idx, clip_range=0.19, vf_coef=0.5, ent_coef=1e-4)
        up.adam(params, 1.5e-4)
    torch.cuda.synchronize()
    assert torch.isfinite(params).all()
print("sanitize ppo done")
