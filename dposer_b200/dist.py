"""Multi-GPU plumbing: one process per GPU, rows sharded contiguously exactly like the reference's
``DistributedEvalSampler`` (lib/dataset/EvaSampler.py:77-106); the only collectives are the result
all-gather and metric all-reduce that replace ``dist.gather_object`` (run/completion.py:300-305).
Backend nccl on GPUs, gloo in the CPU tests."""
import os

import torch
import torch.distributed as dist

from .misc import shard_range


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (no-op for a single process)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


def my_shard(total, units=1):
    """(start, count) of this rank's contiguous rows; ``units`` keeps groups of rows (e.g. 60-frame
    sequences) on one rank by sharding whole units."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    assert total % units == 0
    s, n = shard_range(total // units, world, rank)
    return s * units, n * units


def all_gather_rows(local, total, out=None):
    """Concatenate ragged row shards from all ranks in rank order -> [total, ...] on every rank.

    One ``all_gather_into_tensor`` on a preallocated [world, pad, ...] buffer (no Python list of per-rank tensors);
    when the shards are equal (total % world == 0) the gathered buffer IS the result, otherwise the at-most-one-row
    padding per rank is cut away.  ``out`` ([world*pad, ...]) may be passed to reuse the buffer across calls."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    counts = [shard_range(total, world, r)[1] for r in range(world)]
    pad = max(counts)
    tail = tuple(local.shape[1:])
    if out is None:
        out = local.new_empty((world * pad,) + tail)
    if local.shape[0] == pad:
        src = local.contiguous()
    else:
        src = local.new_zeros((pad,) + tail)
        src[:local.shape[0]] = local
    dist.all_gather_into_tensor(out, src)
    if total % world == 0:
        return out
    return torch.cat([out[r * pad:r * pad + c] for r, c in enumerate(counts)], dim=0)


def all_reduce_sum(t):
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t


def all_reduce_mean_(t):
    """In-place mean over ranks: the data-parallel gradient all-reduce of the training step (one flat buffer, one
    collective).  NCCL averages inside the collective; gloo (CPU tests) sums and divides."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == 'nccl':
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(t)
            t.div_(dist.get_world_size())
    return t


def max_over_ranks(value, device):
    """Max of a python float over ranks (timing rule: report the slowest rank)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
