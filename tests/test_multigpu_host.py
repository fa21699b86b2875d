"""Host-side logic of the slab decomposition, on CPU with gloo and world_size 2 (and the pure partition arithmetic):
after `exchange_halos` every rank's halo rows must equal the neighbouring rows of the global array, for a periodic
ring and for physical (non-periodic) ends."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spruce_b200.multigpu import HALO, exchange_halos, neighbours, partition


def test_partition_covers_all_rows():
    for xdim, world in [(4096, 8), (16384, 8), (97, 2), (33, 4), (16, 4)]:
        parts = partition(xdim, world)
        assert parts[0][0] == 0 and sum(n for _, n in parts) == xdim
        assert all(parts[k][0] + parts[k][1] == parts[k + 1][0] for k in range(world - 1))
        assert max(n for _, n in parts) - min(n for _, n in parts) <= 1
    with pytest.raises(ValueError):
        partition(12, 4)


def test_neighbours():
    assert neighbours(0, 4, True) == (3, 1) and neighbours(3, 4, True) == (2, 0)
    assert neighbours(0, 4, False) == (None, 1) and neighbours(3, 4, False) == (2, None)
    assert neighbours(0, 1, False) == (None, None)


def _worker(rank, world, port, periodic, xdim, ydim, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        G = rng.standard_normal((3, xdim, ydim))            # 3 "planes"
        row0, n = partition(xdim, world)[rank]
        loc = G[:, row0:row0 + n]
        send_lo = torch.from_numpy(np.ascontiguousarray(loc[:, :HALO]).ravel().copy())
        send_hi = torch.from_numpy(np.ascontiguousarray(loc[:, n - HALO:]).ravel().copy())
        recv_lo, recv_hi = torch.full_like(send_lo, np.nan), torch.full_like(send_hi, np.nan)
        exchange_halos(dist, send_lo, send_hi, recv_lo, recv_hi, rank, world, periodic)
        lo, hi = neighbours(rank, world, periodic)
        ok = True
        if lo is not None:
            want = G[:, [(row0 - 2) % xdim, (row0 - 1) % xdim]]
            ok &= np.array_equal(recv_lo.numpy().reshape(3, HALO, ydim), want)
        else:
            ok &= bool(torch.isnan(recv_lo).all())
        if hi is not None:
            want = G[:, [(row0 + n) % xdim, (row0 + n + 1) % xdim]]
            ok &= np.array_equal(recv_hi.numpy().reshape(3, HALO, ydim), want)
        else:
            ok &= bool(torch.isnan(recv_hi).all())
        # the dt minimum is an all-reduce(min) of one double
        t = torch.tensor([float(rank + 1.5)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok &= (t.item() == 1.5)
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("periodic", [True, False])
def test_halo_exchange_gloo_world2(periodic):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), periodic, 21, 10, ret), nprocs=world, join=True)
    assert all(ret.get(r) is True for r in range(world)), dict(ret)
