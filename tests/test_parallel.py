"""Multi-rank host logic on CPU (gloo, world_size 2): env sharding, the flat gradient all-reduce and the
episode-statistics reduction.  The kernels themselves need a GPU; what is checked here is that the N>1
plumbing around them is right and that sharding by global env id leaves results rank-count independent."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import philox
from oracle.envs import OracleEnv
from uav_reinforcement_learning_control_b200 import config as Q
from uav_reinforcement_learning_control_b200 import model as M
from uav_reinforcement_learning_control_b200.parallel import flat_allreduce_mean_, reduce_stats, shard_range

from .util import HostHarness, make_planes, planes_view


def test_shard_range_partitions_exactly():
    for total in (1, 7, 8, 1 << 20, (1 << 23) + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_sharded_reset_equals_single_shard():
    """Philox is keyed by the GLOBAL env id: two shards of 500 reproduce one shard of 1000 bit for bit."""
    n = 1000
    whole = make_planes(n)
    HostHarness(Q.EnvConfig.north_star(seed=5)).reset(whole)
    parts = []
    for r in range(2):
        off, cnt = shard_range(n, 2, r)
        st = make_planes(cnt)
        HostHarness(Q.EnvConfig.north_star(seed=5, env_id_offset=off)).reset(st)
        parts.append(st)
    np.testing.assert_array_equal(np.concatenate(parts, axis=1).view(np.uint32), whole.view(np.uint32))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.ones(12, 128), torch.ones(128), torch.ones(128, 4)]
    grads = [torch.full_like(p, float(rank + 1)) * (i + 1) for i, p in enumerate(params)]
    grads.append(None)
    flat_allreduce_mean_(grads, world)
    ok = all(torch.allclose(g, torch.full_like(g, 1.5 * (i + 1))) for i, g in enumerate(grads[:3]))
    stats = reduce_stats({"episodes": 10 * (rank + 1), "reward_sum": 2.5, "steps": 100}, world)
    # sharded env stepping: each rank steps ITS shard; rank 0 gathers and compares with the unsharded run
    n = 64
    off, cnt = shard_range(n, world, rank)
    cfg = Q.EnvConfig.north_star(seed=9, env_id_offset=off)
    hh = HostHarness(cfg)
    st = make_planes(cnt)
    hh.reset(st)
    ids = np.arange(cnt, dtype=np.uint32) + np.uint32(off)
    for t in range(5):
        raw = philox.draw_blocks(9, ids, np.uint32(t), 1, philox.STREAM_ACTION)
        act = np.stack([philox.uniform(raw[:, i], -1.0, 1.0) for i in range(4)], axis=1)
        hh.step(st, act)
    gathered = [torch.zeros(Q.NPLANES, n // world) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(st.copy()))
    if rank == 0:
        out.put((ok, stats, torch.cat(gathered, dim=1).numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_and_sharded_stepping():
    world, port = 2, 29531 + os.getpid() % 500
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    ok, stats, sharded = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
    assert stats == {"episodes": 30.0, "reward_sum": 5.0, "steps": 200.0}
    # the same 64 envs stepped as ONE shard
    n = 64
    hh = HostHarness(Q.EnvConfig.north_star(seed=9))
    st = make_planes(n)
    hh.reset(st)
    ids = np.arange(n, dtype=np.uint32)
    for t in range(5):
        raw = philox.draw_blocks(9, ids, np.uint32(t), 1, philox.STREAM_ACTION)
        act = np.stack([philox.uniform(raw[:, i], -1.0, 1.0) for i in range(4)], axis=1)
        hh.step(st, act)
    np.testing.assert_array_equal(sharded.view(np.uint32)[:27], st.view(np.uint32)[:27])
