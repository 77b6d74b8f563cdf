"""Sharding of independent ciphertexts / ring elements across the GPUs of one box.

The path has no exchange step: every unit (a ring element in one limb for CRT, a whole ciphertext for
mul+key-switch and tunnelling -- never a single limb of one, because modSwitch and gadget decomposition
couple the limbs of a coefficient) is independent, plans and hints are replicated, and the only collective
is the gather of results ("NCCL over NVLink only to gather results", BASELINE.json).  One process per GPU;
`torch.distributed` (nccl on GPUs, gloo in the CPU tests) is the plumbing.
"""
import torch
import torch.distributed as dist


def shard_range(total, world, rank):
    """contiguous block [lo, hi) of `total` units owned by `rank`; sizes differ by at most one"""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(total, world):
    return [shard_range(total, world, r)[1] - shard_range(total, world, r)[0] for r in range(world)]


def scatter_units(full, world=None, rank=None):
    """this rank's contiguous slice of a batch-leading tensor that every rank can index (replicated input)"""
    world = dist.get_world_size() if world is None else world
    rank = dist.get_rank() if rank is None else rank
    lo, hi = shard_range(full.shape[0], world, rank)
    return full[lo:hi]


def gather_units(local, total, dst=None, packed=False):
    """concatenate every rank's result along the batch axis (ragged shards allowed).
    dst=None: all ranks get the result (all_gather); dst=r: only rank r (others get None).
    packed=True (CUDA int64 residues only): every shard travels in the wire format (alchemy_b200/wire.py: one uint32 per
    residue plus a checksummed header) -- half the NVLink bytes of the int64 tensors; the receiver verifies each shard."""
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = shard_sizes(total, world)
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    pad = max(sizes)
    buf = local
    if local.shape[0] != pad:       # pad ragged shards so one fixed-size collective suffices
        buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf[:local.shape[0]] = local
    buf = buf.contiguous()
    rest = 1
    for d in local.shape[1:]:
        rest *= int(d)
    if packed and local.is_cuda and local.dtype == torch.int64 and 0 < rest < (1 << 24):
        from . import wire
        w = wire.pack("elements", 1, buf.reshape(pad, 1, 1, rest), [])
        if dst is None:
            got = [torch.empty_like(w) for _ in range(world)]
            dist.all_gather(got, w)
        else:
            got = [torch.empty_like(w) for _ in range(world)] if rank == dst else None
            dist.gather(w, got, dst=dst)
            if rank != dst:
                return None
        parts = [wire.unpack(g)[1].reshape((pad,) + tuple(local.shape[1:])) for g in got]
        return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
    if dst is None:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
    else:
        parts = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
        dist.gather(buf, parts, dst=dst)
        if rank != dst:
            return None
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
