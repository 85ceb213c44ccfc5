"""Times one PPO minibatch update (configs[2] shape: 8192 envs x 1024 steps, 8 minibatches of 2^20 samples):
fused sm_100a kernels (qs_ppo_grad + qs_ppo_adam) against the torch-autograd update they replace.  CUDA events."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from uav_reinforcement_learning_control_b200.ppo import ActorCritic, FusedUpdater
    N = int(os.environ.get("PPO_N", 8192 * 1024)); mb = N // 8
    dev = "cuda:0"
    g = torch.Generator(device=dev); g.manual_seed(0)
    obs = torch.rand(N, 12, device=dev, generator=g) * 2 - 1
    act = torch.randn(N, 4, device=dev, generator=g) * 0.4
    old_logp = torch.randn(N, device=dev, generator=g) * 0.1 - 1.0
    adv = torch.randn(N, device=dev, generator=g); ret = torch.randn(N, device=dev, generator=g)
    ac = ActorCritic(12, dev, seed=0, log_std_init=-1.0)
    params = ac.pack()
    up = FusedUpdater(dev)
    perm = torch.randperm(N, device=dev, generator=g).to(torch.int32)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    out = {"samples_per_minibatch": mb}

    packed = up.pack(obs, act, old_logp, adv, ret)
    use_packed = os.environ.get("PPO_PACKED", "1") != "0"

    def fused(k, pk=None):
        pk = use_packed if pk is None else pk
        up.grad(params, obs, act, old_logp, adv, ret, idx=perm[k * mb:(k + 1) * mb], clip_range=0.19, vf_coef=0.5, ent_coef=1e-4,
                packed=packed if pk else None)
        up.adam(params, 1.5e-4)
    for mode in (() if os.environ.get("PPO_ONLY_FUSED") else (False, True, False, True)):
        for k in range(3):
            fused(k, mode)
        torch.cuda.synchronize()
        ev[0].record()
        for rep in range(4):
            for k in range(8):
                fused(k, mode)
        ev[1].record(); torch.cuda.synchronize()
        out.setdefault("packed_rows_ms_per_minibatch" if mode else "separate_arrays_ms_per_minibatch", []).append(ev[0].elapsed_time(ev[1]) / 32)
    if os.environ.get("PPO_ONLY_FUSED"):
        # profiling mode (ncu): a few launches of the selected gather mode only
        for k in range(4):
            fused(k)
        torch.cuda.synchronize()
        print(json.dumps(out))
        return
    ev[0].record()
    for k in range(8):
        fused(k)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 8
    out["fused_ms_per_minibatch"] = ms
    out["fused_samples_per_s"] = mb / ms * 1e3
    out["fused_tflops_bf16"] = mb * 2 * 110336 / ms / 1e9      # 3 x (18 432 + 18 048 + ...) MACs: fwd + dX + dW GEMMs

    opt = torch.optim.Adam(ac.parameters(), lr=1.5e-4, eps=1e-5)

    def autograd(k):
        idx = perm[k * mb:(k + 1) * mb].long()
        a = adv[idx]; a = (a - a.mean()) / (a.std() + 1e-8)
        logp, value, ent = ac.evaluate(obs[idx], act[idx])
        ratio = torch.exp(logp - old_logp[idx])
        pg = torch.max(-a * ratio, -a * torch.clamp(ratio, 0.81, 1.19)).mean()
        loss = pg + 0.5 * torch.nn.functional.mse_loss(value, ret[idx]) - 1e-4 * ent
        opt.zero_grad(set_to_none=True); loss.backward()
        torch.nn.utils.clip_grad_norm_(ac.parameters(), 0.5); opt.step()
    for k in range(2):
        autograd(k)
    torch.cuda.synchronize()
    ev[2].record()
    for k in range(8):
        autograd(k)
    ev[3].record(); torch.cuda.synchronize()
    ms2 = ev[2].elapsed_time(ev[3]) / 8
    out["autograd_ms_per_minibatch"] = ms2
    out["speedup"] = ms2 / ms
    print(json.dumps(out))


if __name__ == "__main__":
    main()
