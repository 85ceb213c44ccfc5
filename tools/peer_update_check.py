"""torchrun, N ranks: the peer-memory update (qs_ppo_adam_peer) against the NCCL all-reduce form of the same minibatch updates:
parameters must agree between the two forms and be bitwise identical across ranks; reports ms per minibatch of both."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from uav_reinforcement_learning_control_b200.parallel import DistContext
from uav_reinforcement_learning_control_b200.ppo import ActorCritic, FusedUpdater

ctx = DistContext.from_env("nccl")
torch.cuda.set_device(ctx.local_rank)
dev = torch.device("cuda", ctx.local_rank)
N = int(os.environ.get("PPO_N", 8192 * 1024)); mb = N // 8
g = torch.Generator(device=dev); g.manual_seed(100 + ctx.rank)            # every rank has its own rollout
obs = torch.rand(N, 12, device=dev, generator=g) * 2 - 1
act = torch.randn(N, 4, device=dev, generator=g) * 0.4
old_logp = torch.randn(N, device=dev, generator=g) * 0.1 - 1.0
adv = torch.randn(N, device=dev, generator=g); ret = torch.randn(N, device=dev, generator=g)
p0 = ActorCritic(12, dev, seed=0, log_std_init=-1.0).pack()
nccl, peer = FusedUpdater(dev), FusedUpdater(dev)
peer.enable_peer(ctx.world, ctx.rank)
perm = nccl.permutation(N, 7 + ctx.rank, 0)
pa, pb = p0.clone(), p0.clone()
stats = torch.zeros(8, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]


def run_nccl(k):
    gr = nccl.grad(pa, obs, act, old_logp, adv, ret, idx=perm[k * mb:(k + 1) * mb], clip_range=0.19, vf_coef=0.5, ent_coef=1e-4)
    dist.all_reduce(gr)
    nccl.adam(pa, 1.5e-4, grad_scale=1.0 / ctx.world)


def run_peer(k):
    peer.grad(pb, obs, act, old_logp, adv, ret, idx=perm[k * mb:(k + 1) * mb], clip_range=0.19, vf_coef=0.5, ent_coef=1e-4)
    peer.adam_peer(pb, 1.5e-4, stats_acc=stats)


for k in range(8):
    run_nccl(k); run_peer(k)
torch.cuda.synchronize(); dist.barrier()
same = float((pa - pb).abs().max()); moved = float((pa - p0).abs().max())
allp = [torch.empty_like(pb) for _ in range(ctx.world)]
dist.all_gather(allp, pb)
in_sync = all(torch.equal(allp[0], t) for t in allp)
out = {}
for name, fn, e0, e1 in (("nccl", run_nccl, ev[0], ev[1]), ("peer", run_peer, ev[2], ev[3])):
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for rep in range(4):
        for k in range(8):
            fn(k)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 32], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[name + "_ms_per_minibatch"] = float(t)
if ctx.rank == 0:
    out.update({"world": ctx.world, "max_abs_diff_peer_vs_nccl": same, "max_abs_param_change": moved,
                "peer_params_bitwise_in_sync_across_ranks": in_sync, "samples_seen_stat": float(stats[4])})
    print(json.dumps(out), flush=True)
peer.close()
dist.barrier(); dist.destroy_process_group()
