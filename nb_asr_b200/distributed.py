"""One-process-per-GPU plumbing (torch.distributed; NCCL on the GPUs, gloo in the CPU tests).

Two ways the path shards:
  * data-parallel training: utterance-sharded batches, ONE exchange per step -- the flat fp32 gradient buffer
    is summed with a single all-reduce and averaged (equal shard sizes => mean of shard means = global mean of
    trainer.py:36-44); the regulariser gradient is identical on every rank so it is added after the reduce.
    Replaces the reference's single-process nn.DataParallel (trainer.py:91-92).
  * architecture-sharded sweep evaluation: independent candidates, no collective on the data path; rows are
    gathered on rank 0 at the end.
"""
import torch
import torch.distributed as dist

from .model import CELLS_PER_BLOCK, CONV_EDGES, FILTERS, HIDDEN
from .search_space import all_ops


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allreduce_mean_(flat):
    """In-place mean over ranks of a flat gradient buffer (no-op for a single process)."""
    _, n = world()
    if n > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.mul_(1.0 / n)
    return flat


class _Done:
    def wait(self):
        return True


def allreduce_mean_async(flat):
    """Start the in-place mean over ranks of one gradient bucket and return a handle whose ``wait()`` orders the CURRENT
    stream after the collective (NCCL: the reduce runs on the process group's own stream, behind the work already queued on
    the current stream, so it overlaps whatever is launched next).  gloo has no AVG: sum, then scale."""
    _, n = world()
    if n == 1:
        return _Done()
    if dist.get_backend() == 'nccl':
        return dist.all_reduce(flat, op=dist.ReduceOp.AVG, async_op=True)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.mul_(1.0 / n)
    return _Done()


def arch_cost(arch_vec, frames=500):
    """Forward FLOPs per utterance: used to balance the sweep (linear edges are ~6x a conv edge)."""
    T = [frames, frames, (frames + 1) // 2, ((frames + 1) // 2 + 1) // 2]
    cin = [80] + FILTERS[:-1]
    total = 0.0
    for b in range(4):
        c = FILTERS[b]
        total += 2.0 * 8 * T[b] * cin[b] * c
        per_cell = 0.0
        for node in arch_vec:
            op = all_ops[node[0]]
            if op == 'linear':
                per_cell += 2.0 * c * c
            elif op in CONV_EDGES:
                per_cell += 2.0 * c * (c // 100) * CONV_EDGES[op][0]
        total += CELLS_PER_BLOCK[b] * T[b] * per_cell
    total += 2.0 * T[3] * 4 * HIDDEN * (FILTERS[-1] + HIDDEN) + 2.0 * T[3] * HIDDEN * 49
    return total


def shard_archs(archs, rank, world_size, balance='lpt', frames=500):
    """Indices of `archs` evaluated by `rank`.  'rr': arch i -> rank i mod N;
    'lpt': longest-processing-time-first on the FLOP model, deterministic on every rank."""
    n = len(archs)
    if balance == 'rr' or world_size == 1:
        return list(range(rank, n, world_size))
    order = sorted(range(n), key=lambda i: (-arch_cost(archs[i], frames), i))
    loads = [0.0] * world_size
    mine = []
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        loads[r] += arch_cost(archs[i], frames)
        if r == rank:
            mine.append(i)
    return sorted(mine)


def gather_rows(rows):
    """Rank 0 receives the concatenation (in rank order) of every rank's list of result rows; others get []."""
    rank, n = world()
    if n == 1:
        return list(rows)
    out = [None] * n if rank == 0 else None
    dist.gather_object(list(rows), out, dst=0)
    if rank != 0:
        return []
    merged = []
    for part in out:
        merged.extend(part)
    return merged
