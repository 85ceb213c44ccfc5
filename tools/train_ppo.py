"""Data-parallel PPO on the fused rollout kernels: one process per GPU (torchrun), envs sharded by global id,
NCCL used only for the flat gradient all-reduce and the episode-statistics reduction.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_ppo.py \
        --total-envs 65536 --iters 10
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total-envs", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--n-steps", type=int, default=64)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--log-json", default=None, help="write the per-iteration log (mean reward, losses, clip fraction) here")
    ap.add_argument("--nccl", action="store_true", help="exchange gradients with an NCCL all-reduce instead of the peer-memory kernel")
    ap.add_argument("--autograd", action="store_true", help="torch-autograd update instead of the fused qs_ppo_grad / qs_ppo_adam kernels")
    args = ap.parse_args()
    import torch
    from uav_reinforcement_learning_control_b200 import config as Q
    from uav_reinforcement_learning_control_b200.engine import Engine
    from uav_reinforcement_learning_control_b200.parallel import DistContext, shard_range
    from uav_reinforcement_learning_control_b200.ppo import PPOConfig, PPOTrainer
    ctx = DistContext.from_env("nccl")
    torch.cuda.set_device(ctx.local_rank)
    off, cnt = shard_range(args.total_envs, ctx.world, ctx.rank)
    eng = Engine(Q.EnvConfig.north_star(seed=0, env_id_offset=off), cnt, device=ctx.local_rank)
    tr = PPOTrainer(eng, PPOConfig(n_steps=args.n_steps, learning_rate=args.lr, ent_coef=0.0), ctx=ctx, seed=0,
                    fused=not args.autograd, peer=not args.nccl)
    tr.set_log_std(-1.0)
    torch.manual_seed(1234)                       # same minibatch permutation stream on every rank
    torch.cuda.synchronize()
    t0 = time.time()
    log = tr.train(args.iters)
    torch.cuda.synchronize()
    dt = time.time() - t0
    # all ranks must hold identical parameters after data-parallel training
    flat = tr.packed_params()
    chk = torch.stack([flat.double().sum(), flat.double().abs().sum()])
    if ctx.world > 1:
        import torch.distributed as dist
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        in_sync = bool(torch.allclose(lo, hi, rtol=0, atol=0))
    else:
        in_sync = True
    if ctx.rank == 0 and args.log_json:
        with open(args.log_json, "w") as f:
            json.dump({"total_envs": args.total_envs, "n_steps": args.n_steps, "iters": args.iters, "seconds": dt,
                       "env_steps": args.total_envs * args.n_steps * args.iters, "log": log}, f, indent=1)
    if ctx.rank == 0:
        print(json.dumps({"world": ctx.world, "update": "fused tcgen05 kernels" if tr.fused else "torch autograd", "total_envs": args.total_envs, "iters": args.iters,
                          "env_steps_per_s_incl_update": args.total_envs * args.n_steps * args.iters / dt,
                          "mean_reward_first": log[0]["mean_reward"], "mean_reward_last": log[-1]["mean_reward"],
                          "episodes_last": log[-1]["episodes"], "params_in_sync_across_ranks": in_sync}), flush=True)
    if tr.updater is not None:
        tr.updater.close()                     # collective tear-down of the peer buffers (unmap, barrier, free)
    if ctx.world > 1:
        import torch.distributed as dist
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
