"""world_size-2 gloo test of the multi-GPU host logic (sharding by email + the one all-gather of
result records); the same code runs over NCCL on the GPU box (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zkemail_rs_b200 import shard
    from zkemail_rs_b200.engine import RESULT_DTYPE
    lo, hi = shard.shard_range(n, rank, world)
    # each rank "verifies" its shard: synthetic records whose content depends on the global index
    local = np.zeros(hi - lo, dtype=RESULT_DTYPE)
    idx = np.arange(lo, hi)
    local["status"] = np.where(idx % 7 == 0, 3, 0)
    local["body_hash"][:, 0] = idx % 251
    full = shard.all_gather_records(local, n, rank, world)
    ok = (len(full) == n and np.array_equal(full["status"], np.where(np.arange(n) % 7 == 0, 3, 0))
          and np.array_equal(full["body_hash"][:, 0], np.arange(n) % 251))
    bits = shard.pack_verdicts(full["status"])
    ok = ok and np.array_equal(np.unpackbits(bits, bitorder="little")[:n].astype(bool), np.arange(n) % 7 != 0)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _worker_api(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zkemail_rs_b200 import shard
    from zkemail_rs_b200.engine import RESULT_DTYPE

    class StubEngine:   # stands in for the per-rank CUDA engine: the record content encodes the email
        def verify_batch(self, emails):
            out = np.zeros(len(emails), dtype=RESULT_DTYPE)
            for i, e in enumerate(emails):
                out["status"][i] = 0 if e % 5 else 3
                out["header_hash"][i, 0] = e % 251
            return out
    emails = list(range(n))
    full = shard.verify_batch_sharded(StubEngine(), emails)
    ok = (len(full) == n and all(int(full["status"][i]) == (0 if i % 5 else 3) for i in range(n))
          and all(int(full["header_hash"][i, 0]) == i % 251 for i in range(n)))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_verify_batch_sharded_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_api, args=(r, 2, port, 777, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


@pytest.mark.parametrize("n", [1001, 2])
def test_shard_and_allgather_world2(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_ranges_cover_and_balance():
    from zkemail_rs_b200 import shard
    for n in (0, 1, 7, 1000, 1_000_003):
        for w in (1, 2, 4, 8):
            rs = [shard.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
    cost = np.concatenate([np.full(1000, 1.0), np.full(1000, 16.0)])
    rs = shard.balanced_ranges(cost, 4)
    sums = [cost[a:b].sum() for a, b in rs]
    assert rs[0][0] == 0 and rs[-1][1] == 2000 and max(sums) / (sum(sums) / 4) < 1.05
