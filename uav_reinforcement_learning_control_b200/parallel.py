"""Multi-GPU plumbing: env sharding (no data-path collective) + the two small collectives PPO needs.

The reference never configures multi-device training itself; Brax's trainer does it internally with
``pmap`` + ``lax.pmean`` of the gradients (third party; train_brax_ppo.py:589-620 only passes
``num_envs``).  The B200 design: envs are independent, so rank r owns the contiguous
global env ids [offset_r, offset_r + count_r) and steps them with ZERO per-step communication; the
Philox streams are keyed by the GLOBAL env id, so any sharding reproduces the single-GPU run bit for
bit.  NCCL is used for exactly two things: one all-reduce of the flattened ~37.5 k-float PPO gradient
per minibatch, and one all-reduce of a handful of episode statistics per rollout.  Both are latency
bound (<= 150 KB); fusing the gradient reduction, a peer-memory all-reduce and the Adam step into one kernel is the next
multi-GPU step.
"""
from __future__ import annotations

from dataclasses import dataclass

__all__ = ["shard_range", "DistContext", "flat_allreduce_mean_", "reduce_stats"]


def shard_range(total_envs: int, world: int, rank: int):
    """Contiguous, balanced split of [0, total_envs): returns (offset, count) of `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(total_envs), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


@dataclass
class DistContext:
    rank: int = 0
    world: int = 1
    local_rank: int = 0
    group: object = None

    @staticmethod
    def from_env(backend: str = "nccl"):
        """torchrun-style env (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*), one process per GPU."""
        import os
        import torch
        import torch.distributed as dist
        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("RANK", "0"))
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if world > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            kw = {}
            if backend == "nccl":
                torch.cuda.set_device(local)
                kw["device_id"] = torch.device("cuda", local)
            dist.init_process_group(backend, **kw)
        return DistContext(rank, world, local, None)


def flat_allreduce_mean_(tensors, world: int, group=None):
    """Average a list of same-dtype tensors across ranks with ONE collective (flatten -> all_reduce -> scatter back).

    The PPO gradient of the 2x128 actor-critic is ~150 KB: a single latency-bound NCCL all-reduce over
    NVLink; bucketing it further would only add launches.
    """
    import torch
    import torch.distributed as dist
    tensors = [t for t in tensors if t is not None]
    if world <= 1 or not tensors:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
    o = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[o:o + n].view_as(t))
        o += n


def reduce_stats(values, world: int, group=None, device=None):
    """Sum a small dict of python/torch scalars across ranks with one all-reduce; returns floats."""
    import torch
    import torch.distributed as dist
    keys = sorted(values)
    t = torch.tensor([float(values[k]) for k in keys], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return {k: float(v) for k, v in zip(keys, t.tolist())}
