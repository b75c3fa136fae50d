"""CPU (gloo, world_size 2): the host-side logic of the view-sharded path -- the view partition and the padded
all-gather whose valid rows must be a contiguous, correctly ordered prefix. No compute kernels are called."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mvdetr_b200.sharded import ViewPartition, gather_rows


@pytest.mark.parametrize("views,world,expect", [
    (7, 1, [(0, 7)]),
    (7, 2, [(0, 4), (4, 7)]),
    (7, 4, [(0, 2), (2, 4), (4, 6), (6, 7)]),
    (7, 8, [(i, i + 1) for i in range(7)] + [(7, 7)]),
    (6, 8, [(i, i + 1) for i in range(6)] + [(6, 6), (6, 6)]),
    (6, 4, [(0, 2), (2, 4), (4, 6), (6, 6)]),
])
def test_view_partition(views, world, expect):
    p = ViewPartition(views, world)
    assert [(p.lo(r), p.hi(r)) for r in range(world)] == expect
    assert sum(p.count(r) for r in range(world)) == views
    # padded gather buffer: the valid rows are a prefix <=> every rank before the last non-empty one is full
    act = p.active_ranks()
    assert all(p.count(r) == p.per_rank for r in act[:-1])
    with pytest.raises(ValueError):
        ViewPartition(0, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, views, rows_per_view, C, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        part = ViewPartition(views, world)
        full = torch.arange(views * rows_per_view * C, dtype=torch.float32).view(views * rows_per_view, C)
        local = full[part.lo(rank) * rows_per_view: part.hi(rank) * rows_per_view].clone()
        got = gather_rows(local, part, rows_per_view, rank)
        ok = torch.equal(got, full)
        # in-place form used by the encoder: local rows already sit in the gather buffer
        buf = torch.zeros(world, part.per_rank * rows_per_view, C)
        buf[rank][:local.shape[0]] = local * 2
        got2 = gather_rows(buf[rank][:local.shape[0]], part, rows_per_view, rank, out=buf)
        ok = ok and torch.equal(got2, full * 2) and got2.data_ptr() == buf.data_ptr()
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("views", [7, 6, 1])
def test_gather_rows_gloo_world2(views):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), views, 5, 3, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
