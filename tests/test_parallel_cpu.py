"""CPU tests of the multi-GPU host logic: env sharding and the episode-statistics all-gather,
world_size 2 over gloo (the path bench.py --gpus N takes with NCCL)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ast_sac_b200 import parallel as P


def test_shard_range_partitions_exactly():
    for total in (1, 7, 1000, 100_000, 10_000_019):
        for world in (1, 2, 3, 4, 8):
            edges = [P.shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            for (a, b), (c, d) in zip(edges[:-1], edges[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = P.shard_range(1001, rank, world)
    n = hi - lo
    g = torch.Generator().manual_seed(100 + rank)
    bits = torch.randint(0, 1 << 11, (n,), generator=g, dtype=torch.int32)
    ret = torch.rand(n, generator=g, dtype=torch.float64)
    steps = torch.randint(1, 1600, (n,), generator=g, dtype=torch.int32)
    rl = torch.randint(1, 10, (n,), generator=g, dtype=torch.int32)
    local = P.episode_stats(bits, ret, steps, rl)
    allv = P.gather_stats(local)
    q.put((rank, local.tolist(), allv.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_stats_allgather_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=60) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    locals_ = [r[1] for r in res]
    for r in res:
        assert r[2] == locals_          # every rank sees every rank's record, in rank order
    s = P.summarise(torch.tensor(res[0][2], dtype=torch.float64))
    assert s["episodes"] == 1001
    assert abs(s["env_steps"] - (locals_[0][P.N_EVENTS + 1] + locals_[1][P.N_EVENTS + 1])) < 1e-9
