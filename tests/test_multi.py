"""Multi-GPU path: cost-balanced contiguous shards, one engine per device behind ONE handle
(zkb_multi_*), and the NCCL all-gather of the fixed-size result records issued by the C++ library (zkb_multi_batch_run,
zkb_comm_allgather_records).  CPU: the shard planner, the record layout and a world-size-2 gloo run of the same slot
layout.  GPU: every visible device (the single-GPU box exercises the same code with one shard and a one-rank
communicator; the two-device test needs `gpurun --gpus 2`)."""
import os
import socket
import sys

import numpy as np
import pytest

import oracle
import zkemail_rs_b200 as z
from zkemail_rs_b200.engine import RESULT_DTYPE, EmailViews
from zkemail_rs_b200.structs import CompiledRegex, RegexInfo
from tests.util import NOW, ROOT, assert_records_equal, mixed_emails


def _views(lens, key_lens):
    bufs = [bytes(l) for l in lens]
    keys = [bytes(k) for k in key_lens]
    return EmailViews.from_emails([z.Email("d.example", b, z.PublicKey(k, "rsa")) for b, k in zip(bufs, keys)])


def test_plan_shards_cover_balance_and_edge_cases():
    rng = np.random.default_rng(4)
    lens = np.concatenate([np.full(3000, 5000), rng.integers(1000, 70000, size=3000)])
    v = _views(lens.tolist(), [270] * 6000)
    for resident in (False, True):
        for w in (1, 2, 3, 4, 8):
            b = z.plan_shards(v, w, resident)
            assert b[0] == 0 and b[-1] == 6000 and all(b[i] <= b[i + 1] for i in range(w))
            per_byte, rsa = (8, 18000) if resident else (20, 18000)
            cost = lens * per_byte + rsa + 2000
            sums = [cost[b[i]:b[i + 1]].sum() for i in range(w)]
            assert max(sums) / (sum(sums) / w) < 1.02, (w, sums)
            cnt = [b[i + 1] - b[i] for i in range(w)]
            if w > 1:
                assert cnt[0] > cnt[-1]          # the small messages come first: more of them per shard
    # more shards than emails, and no emails at all: trailing shards are empty
    assert z.plan_shards(_views([10, 10], [270, 270]), 4) == [0, 1, 1, 2, 2] or z.plan_shards(_views([10, 10], [270, 270]), 4)[-1] == 2
    assert z.plan_shards(EmailViews.from_emails([]), 3) == [0, 0, 0, 0]
    # key size enters the cost: 1024-bit keys are cheaper than 2048-bit ones in resident batches
    v2 = _views([1000] * 2000, [140] * 1000 + [270] * 1000)
    b = z.plan_shards(v2, 2, True)
    assert b[1] > 1000


def test_record_layout_is_the_head_of_zkb_result():
    assert RESULT_DTYPE.fields["parts"][1] == 144 and RESULT_DTYPE.itemsize == 400
    recs = np.zeros((3, 144 + 32), dtype=np.uint8)
    recs[:, 0] = [0, 3, 7]
    recs[1, 8:40] = 0xAB
    recs[2, 144:148] = [1, 0, 0, 0]
    r = z.records_to_results(recs, 2)
    assert r["status"].tolist() == [0, 3, 7] and bytes(r["body_hash"][1]) == b"\xab" * 32 and int(r["parts"][2][0][0]) == 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slot_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zkemail_rs_b200 import shard
    import zkemail_rs_b200 as zz
    # ranks hold cost-balanced shards of different sizes of one 900-email batch; records are 144 + 16 * 3 bytes
    lens = [5000] * 600 + [60000] * 300
    bounds = zz.plan_shards(_views(lens, [270] * 900), world, True)
    lo, hi = bounds[rank], bounds[rank + 1]
    rec = np.zeros((hi - lo, 192), dtype=np.uint8)
    idx = np.arange(lo, hi)
    rec[:, 0] = np.where(idx % 7 == 0, 3, 0)
    rec[:, 8] = idx % 251
    rec[:, 144] = idx % 2
    gathered, counts = shard.gather_record_slots(rec.view(np.dtype((np.void, 192))).reshape(-1), rank, world)
    ok = counts == [bounds[r + 1] - bounds[r] for r in range(world)] and gathered.shape == (world, max(counts), 192)
    full = np.concatenate([gathered[r, :counts[r]] for r in range(world)])
    res = zz.records_to_results(full, 3)
    n = 900
    ok = ok and np.array_equal(res["status"], np.where(np.arange(n) % 7 == 0, 3, 0)) and np.array_equal(res["body_hash"][:, 0], np.arange(n) % 251)
    ok = ok and np.array_equal(res["parts"][:, 0, 0], np.arange(n) % 2)
    for r in range(world):      # padding of the shorter slots is zero
        ok = ok and not gathered[r, counts[r]:].any()
    q.put((rank, bool(ok), counts))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_record_slot_allgather_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_slot_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(r[:2] for r in res) == [(r, True) for r in range(world)]
    assert res[0][2][0] > res[0][2][-1]      # unequal shard sizes were exercised


# ------------------------------------------------------------------ GPU
def _device_count():
    import ctypes
    try:
        rt = ctypes.CDLL("libcudart.so.12")
    except OSError:
        rt = ctypes.CDLL("libcudart.so")
    n = ctypes.c_int()
    return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0


def _multi_roundtrip(n_devices):
    emails, labels = mixed_emails(seed=51, n_pos=90, with_token=True)
    emails = emails * 3
    info = RegexInfo([CompiledRegex(z.compile_regex(r"\r\nsubject:[^\r\n]+\r\n"), None)],
                     [CompiledRegex(z.compile_regex(r"Transaction ID: [A-Z0-9]+"), None)])
    exp = oracle.verify_batch(emails, info.header_parts, info.body_parts, now=NOW)
    exp_plain = oracle.verify_batch(emails, now=NOW)
    m = z.MultiEngine(n_devices=n_devices, now_unix=NOW, chunk_emails=64)
    try:
        got = m.verify_batch(emails)                                   # end to end: every device writes its range
        for i, (g, e) in enumerate(zip(got, exp_plain)):
            assert_records_equal(g, e, i)
        rs = z.MultiRegexSet(m, info)
        got = m.verify_views(EmailViews.from_emails(emails), rs)
        for i, (g, e) in enumerate(zip(got, exp)):
            assert_records_equal(g, e, i)
        views = EmailViews.from_emails(emails)
        for raw in (False, True):                                      # resident shards + the all-gather of the records
            mb = m.prepare(views, rs, raw=raw)
            b = mb.bounds()
            assert b[0] == 0 and b[-1] == len(emails) and len(b) == n_devices + 1
            ms = mb.run(gather=True)
            assert ms > 0
            mb.run(gather=True)
            res = mb.fetch()
            for i, (g, e) in enumerate(zip(res, exp)):
                assert_records_equal(g, e, (raw, i))
            for d in range(n_devices):                                 # every device holds every shard's records
                gat = mb.gathered(d)
                for i, (g, r) in enumerate(zip(gat, res)):
                    if int(g["status"]) == 0x7fffffff:
                        continue                                       # declined on the device: finished by the host in fetch
                    assert g.tobytes() == r.tobytes(), (raw, d, i)
                assert sum(int(g["status"]) == 0 for g in gat) >= 0.5 * sum(int(r["status"]) == 0 for r in res)
            mb.close()
        rs.close()
    finally:
        m.close()


@pytest.mark.gpu
def test_multi_engine_on_one_device():
    _multi_roundtrip(1)


@pytest.mark.gpu
def test_multi_engine_on_two_devices():
    if _device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    _multi_roundtrip(2)


@pytest.mark.gpu
@pytest.mark.parametrize("overlapped", [False, True])
def test_comm_allgather_of_resident_records_one_rank(engine, overlapped):
    """zkb_comm with a world of one: the library's NCCL path (id, communicator, count exchange, in-place all-gather) and the
    run + exchange call (chunk table, exchange stream; no peers to send to)."""
    import ctypes as C
    emails, _ = mixed_emails(seed=52, n_pos=40)
    views = EmailViews.from_emails(emails)
    comm = z.Comm(engine, 0, 1)
    try:
        for raw in (False, True):
            pb = engine.prepare(views, raw=raw)
            if overlapped:
                ptr, slot, rb = comm.run_allgather(pb)
            else:
                pb.run_async()
                ptr, slot, rb = comm.allgather_records(pb)
            res = pb.fetch()                      # synchronises the engine stream
            assert slot == len(emails) and rb == 144 and comm.rank_records() == [len(emails)]
            host = np.zeros((slot, rb), dtype=np.uint8)
            rt = C.CDLL("libcudart.so.12")
            assert rt.cudaMemcpy(C.c_void_p(host.ctypes.data), C.c_void_p(ptr), C.c_size_t(host.nbytes), 2) == 0
            gat = z.records_to_results(host, 0)
            for g, r in zip(gat, res):
                if int(g["status"]) != 0x7fffffff:
                    assert g.tobytes()[:144] == r.tobytes()[:144]
            pb.close()
    finally:
        comm.close()


def _comm_worker(rank, world, port, q):
    """One process per GPU: ranks hold different shards in different numbers of chunks; every rank must end up with every
    rank's records, byte for byte what the owner fetched."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    try:
        import ctypes as C
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import zkemail_rs_b200 as zz
        from tests.util import mixed_emails as mk, NOW as now
        emails, _ = mk(seed=60, n_pos=120)
        lo, hi = (0, len(emails) // 3) if rank == 0 else (len(emails) // 3, len(emails))
        eng = zz.Engine(device=rank, now_unix=now, chunk_emails=8)          # resident chunks of 32 emails: 2 and 3 exchange rounds
        comm = zz.Comm(eng, rank, world)
        from zkemail_rs_b200.engine import EmailViews as EV
        views = EV.from_emails(emails[lo:hi])
        ok = True
        rt = C.CDLL("libcudart.so.12")
        for overlapped in (True, False, True):
            pb = eng.prepare(views, raw=True)
            if overlapped:
                ptr, slot, rb = comm.run_allgather(pb)
            else:
                pb.run_async()
                ptr, slot, rb = comm.allgather_records(pb)
            mine = pb.fetch()
            counts = comm.rank_records()
            host = np.zeros((world, slot, rb), dtype=np.uint8)
            ok = ok and rt.cudaSetDevice(rank) == 0
            ok = ok and rt.cudaMemcpy(C.c_void_p(host.ctypes.data), C.c_void_p(ptr), C.c_size_t(host.nbytes), 2) == 0
            own = [bytes(host[rank, i]) for i in range(counts[rank])]
            box = [None] * world
            dist.all_gather_object(box, own)
            for r in range(world):
                ok = ok and counts[r] == len(box[r]) and not host[r, counts[r]:].any()
                for i in range(counts[r]):
                    ok = ok and bytes(host[r, i]) == box[r][i]
            for i, m in enumerate(mine):                                     # and the owner's slot is what fetch() returns
                if int.from_bytes(own[i][:4], "little", signed=True) != 0x7fffffff:
                    ok = ok and own[i] == m.tobytes()[:rb]
            pb.close()
        # an empty shard: rank 0 holds nothing (a batch smaller than the world), rank 1 everything
        part = emails[:0] if rank == 0 else emails[:50]
        pb = eng.prepare(EV.from_emails(part), raw=False)
        ptr, slot, rb = comm.run_allgather(pb)
        pb.fetch()
        counts2 = comm.rank_records()
        ok = ok and counts2 == [0, 50] and slot == 50
        host = np.zeros((world, slot, rb), dtype=np.uint8)
        ok = ok and rt.cudaDeviceSynchronize() == 0     # fetch() of an empty batch has nothing to wait for
        ok = ok and rt.cudaMemcpy(C.c_void_p(host.ctypes.data), C.c_void_p(ptr), C.c_size_t(host.nbytes), 2) == 0
        ok = ok and not host[0].any() and host[1].any()
        pb.close()
        q.put((rank, bool(ok), counts))
        comm.close()
        eng.close()
        dist.destroy_process_group()
    except Exception as ex:   # noqa: BLE001
        q.put((rank, False, repr(ex)))


@pytest.mark.gpu
def test_comm_run_allgather_two_ranks():
    if _device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_comm_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=600) for _ in ps]
    for p in ps:
        p.join(timeout=120)
    assert sorted(r[:2] for r in res) == [(0, True), (1, True)], res
