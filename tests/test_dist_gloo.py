"""CPU: the N>1 plumbing (contiguous sharding, ragged all-gather, metric all-reduce) with the gloo
backend and world_size 2 -- the same code path bench.py uses with nccl."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from dposer_b200 import dist as D
    r, _, w = D.init_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    full = torch.arange(total * 3, dtype=torch.float32).view(total, 3)
    s, n = D.my_shard(total)
    gathered = D.all_gather_rows(full[s:s + n].clone(), total)
    ok = torch.equal(gathered, full)
    # the preallocated-buffer form bench.py's e2e region uses: same result, and when the shards are equal the buffer IS the result
    buf = torch.empty(world * ((total + world - 1) // world), 3)
    g2 = D.all_gather_rows(full[s:s + n].clone(), total, out=buf)
    ok = ok and torch.equal(g2, full) and (total % world != 0 or g2.data_ptr() == buf.data_ptr())
    # whole 60-frame sequences stay on one rank
    s2, n2 = D.my_shard(7 * 60, units=60)
    ok = ok and s2 % 60 == 0 and n2 % 60 == 0
    part = torch.tensor([float(full[s:s + n].sum())], dtype=torch.float64)
    tot = D.all_reduce_sum(part)
    ok = ok and abs(float(tot) - float(full.sum())) < 1e-6
    g = torch.full((5,), float(rank + 1))
    D.all_reduce_mean_(g)                       # the training step's gradient all-reduce (mean over ranks)
    ok = ok and bool((g == (world + 1) / 2).all())
    mx = D.max_over_ranks(10.0 + rank, 'cpu')
    ok = ok and mx == 10.0 + world - 1
    q.put((rank, bool(ok), s, n))
    dist.destroy_process_group()


def test_world_size_2_gloo_shard_gather_reduce():
    ctx = mp.get_context('spawn')
    for total in (13, 500):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=120) for _ in procs)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert all(r[1] for r in res)
        assert res[0][2] == 0 and res[0][3] + res[1][3] == total and res[1][2] == res[0][3]
