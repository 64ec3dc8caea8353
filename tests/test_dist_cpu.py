"""world_size-2 gloo tests (CPU) of the N > 1 host logic: every rank plans its
slice of the pair-tile work list, the slices partition it, and the bench's
max-over-ranks timing reduction works."""
import ctypes
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, out):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pyiid_b200 import _lib
    from pyiid_b200.backend import _dist_state
    import bench
    assert _dist_state() == (rank, world)
    lib = _lib.load()
    types = np.zeros(n, np.int32)
    types[n // 2:] = 1
    res = []
    for tri in (0, 1):
        v = [ctypes.c_int64(0) for _ in range(4)]
        assert lib.iid_plan_shard(n, types.ctypes.data, 2, 148, tri, rank, world,
                                  *[ctypes.byref(x) for x in v]) == 0
        t = torch.tensor([v[1].value, v[2].value], dtype=torch.int64)
        dist.all_reduce(t)
        res += [v[0].value, int(t[0]), int(t[1]), v[3].value]
    # the bench's cross-rank timing reduction: max over ranks
    tmax = bench.max_over_ranks(10.0 + rank)
    units = bench.sum_over_ranks(5.0)
    dist.barrier()
    if rank == 0:
        out.put((res, tmax, units))
    dist.destroy_process_group()


def test_two_rank_shards_partition_the_work():
    world, n = 2, 1500
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, out))
             for r in range(world)]
    for p in procs:
        p.start()
    res, tmax, units = out.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    total_sq, items_sq, slots_sq, npad = res[:4]
    total_tri, items_tri, slots_tri, _ = res[4:]
    assert items_sq == total_sq and items_tri == total_tri
    assert slots_sq == npad * npad
    nt = npad // 32
    assert slots_tri == (nt * (nt - 1) // 2 + nt) * 1024
    assert tmax == 11.0 and units == 10.0
