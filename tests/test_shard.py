"""Multi-GPU host logic on CPU: world_size-2 gloo processes shard independent units and gather results."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alchemy_b200 import shard


def test_shard_ranges_cover_and_balance():
    for total in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            rs = [shard.shard_range(total, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1 and sizes == shard.shard_sizes(total, world)


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(total * 6, dtype=torch.int64).reshape(total, 2, 3)
        mine = shard.scatter_units(full)
        # stand-in for the per-GPU hot path: any per-unit map commutes with sharding
        res = (mine * 3 + 1) % 1000003
        got_all = shard.gather_units(res, total)
        got_dst = shard.gather_units(res, total, dst=0)
        want = (full * 3 + 1) % 1000003
        ok = torch.equal(got_all, want) and ((got_dst is None) if rank else torch.equal(got_dst, want))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7, 1])
def test_world_size_2_gloo_scatter_compute_gather(total):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert results == [(0, True), (1, True)]


def _nccl_worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        g = torch.Generator().manual_seed(3)
        full = torch.randint(0, 1 << 31, (total, 2, 3, 64), dtype=torch.int64, generator=g).cuda()
        mine = shard.scatter_units(full).clone()
        ok = torch.equal(shard.gather_units(mine, total, packed=True), full)
        at0 = shard.gather_units(mine, total, dst=0, packed=True)
        ok = ok and ((at0 is None) if rank else torch.equal(at0, full))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("total", [9, 2])
def test_packed_gather_over_nccl(total):
    """results gathered in the wire format (uint32 payload, checksummed): as many ranks as there are GPUs, at most two"""
    world = min(2, torch.cuda.device_count())
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert results == [(r, True) for r in range(world)]
