"""CPU, world_size 2, gloo: host-side logic of the multi-GPU paths (gradient mean all-reduce, sweep sharding, gather)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nb_asr_b200 import distributed as D
from nb_asr_b200 import search_space as ss


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # data-parallel gradient exchange: flat buffer -> mean over ranks
        g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        D.allreduce_mean_(g)
        ok_grad = torch.allclose(g, torch.arange(1000, dtype=torch.float32) * (1 + world) / 2)
        # bucketed exchange (trainer._step_graph): disjoint ranges of one flat buffer, each reduced on its own, waited at the end
        g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        works = [D.allreduce_mean_async(g[lo:hi]) for lo, hi in ((640, 1000), (192, 640), (0, 192))]
        for w in works:
            w.wait()
        ok_grad = ok_grad and torch.allclose(g, torch.arange(1000, dtype=torch.float32) * (1 + world) / 2)
        # sweep sharding: a partition of the arch list, identical on every rank, balanced by the FLOP model
        archs = list(ss.get_all_architectures())[:97]
        mine = D.shard_archs(archs, rank, world, balance='lpt')
        rows = D.gather_rows([dict(index=i, cost=D.arch_cost(archs[i])) for i in mine])
        if rank == 0:
            idx = sorted(r['index'] for r in rows)
            loads = [sum(D.arch_cost(archs[i]) for i in D.shard_archs(archs, r, world)) for r in range(world)]
            q.put((ok_grad, idx == list(range(97)), max(loads) / min(loads)))
        else:
            q.put((ok_grad, True, 1.0))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[0] for r in res)                  # gradient mean is exact on both ranks
    assert all(r[1] for r in res)                  # shards partition the list, gather restores all rows
    assert max(r[2] for r in res) < 1.05           # LPT balance within 5 %


def test_shard_rr_and_single_process():
    archs = list(ss.get_all_architectures())[:10]
    assert D.shard_archs(archs, 1, 4, balance='rr') == [1, 5, 9]
    assert D.shard_archs(archs, 0, 1) == list(range(10))
    assert D.gather_rows([1, 2]) == [1, 2]
    lin = D.arch_cost([[0, 1], [0, 1, 1], [0, 1, 1, 1]])
    c7 = D.arch_cost([[4, 1], [4, 1, 1], [4, 1, 1, 1]])
    assert abs(lin / 1e9 - 35.58) < 0.05 and abs(c7 / 1e9 - 12.42) < 0.05     # SURVEY.md §8d GFLOP / utterance
