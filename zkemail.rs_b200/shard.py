"""Multi-GPU plumbing for the batch path: the batch shards by email (no cross-email dependency
exists anywhere in core/src/circuits.rs:9-68), one process per GPU, and the ONLY exchange is an
all-gather of the fixed-size per-email verdict records / bitmap.  Works on any
torch.distributed backend: NCCL over NVLink on the GPU box, gloo in the CPU tests."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous email range [lo, hi) of `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balanced_ranges(cost: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous ranges with (nearly) equal summed cost, for mixed body sizes / key sizes
    (cost ~ sha_blocks * 1400 + rsa_macs per email)."""
    n = len(cost)
    if n == 0:
        return [(0, 0)] * world
    c = np.cumsum(cost, dtype=np.float64)
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(c, c[-1] * r / world, side="left")))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.array(cuts))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def pack_verdicts(status: np.ndarray) -> np.ndarray:
    """status (int32 per email) -> bitmap, bit i = email i verified (status == 0)."""
    return np.packbits(status == 0, bitorder="little")


def all_gather_records(local: np.ndarray, n_total: int, rank: int, world: int, device=None):
    """All-gathers per-email fixed-size records (any numpy dtype) so that every rank holds the
    records of the whole batch in email order.  Shards are padded to the largest shard (the
    collective needs equal sizes) and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    width = local.dtype.itemsize
    mx = max(sizes)
    buf = np.zeros((mx, width), dtype=np.uint8)
    buf[: len(local)] = local.view(np.uint8).reshape(len(local), width)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    parts = [o.cpu().numpy()[: sizes[r]] for r, o in enumerate(outs)]
    return np.concatenate(parts, axis=0).reshape(-1).view(local.dtype)


def verify_batch_sharded(engine, emails, regex_info=None, device=None):
    """verify_email over a batch sharded by email across the ranks of the default process group: every
    rank verifies its contiguous range on its own GPU (engine = this rank's Engine) and one all-gather of
    the fixed-size result records gives every rank the records of the whole batch, in email order.
    Works with one process (no process group) as a plain batch call."""
    import torch.distributed as dist
    n = len(emails)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return engine.verify_with_regex_batch(emails, regex_info) if regex_info is not None else engine.verify_batch(emails)
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(n, rank, world)
    mine = emails[lo:hi]
    local = engine.verify_with_regex_batch(mine, regex_info) if regex_info is not None else engine.verify_batch(mine)
    return all_gather_records(local, n, rank, world, device=device)


def gather_record_slots(local: np.ndarray, rank: int, world: int, device=None):
    """The layout of zkb_comm_allgather_records (csrc/engine_multi.inc) on any torch.distributed backend: the per-rank
    record counts are exchanged first, every rank's records go into a slot of `slot` = max count records (zero padded),
    slots are all-gathered; returns (gathered [world, slot, width] uint8, counts).  Used by the gloo tests of the
    multi-GPU host logic; on the GPU box the library issues the same exchange with NCCL itself."""
    import torch
    import torch.distributed as dist
    width = local.dtype.itemsize
    cnt = torch.tensor([len(local)], dtype=torch.int64)
    if device is not None:
        cnt = cnt.to(device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    counts = [int(c.item()) for c in cnts]
    slot = max(counts)
    buf = np.zeros((slot, width), dtype=np.uint8)
    buf[: len(local)] = local.view(np.uint8).reshape(len(local), width)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return np.stack([o.cpu().numpy() for o in outs]), counts
