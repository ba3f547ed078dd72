"""Multi-GPU bindings (include/zkemail_b200.h, "multi-GPU" section).

The batch shards by email — verify_email / verify_email_with_regex are pure per-email functions
(core/src/circuits.rs:9,31) — into contiguous cost-balanced ranges, one engine per device.  The only exchange is an NCCL
all-gather of the fixed-size result records, issued by the C++ library:

  * MultiEngine  : ONE process, one engine + host thread per device (zkb_multi_*);
  * Comm         : one process per device (torchrun); rank 0 makes the 128-byte id, the caller broadcasts it (here:
                   torch.distributed, any backend), the collective itself runs in the library (zkb_comm_*).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from .engine import RESULT_DTYPE, EmailViews, Engine, PreparedBatch, _check, _DfaView, _Options, load_library
from .structs import Email, RegexInfo

COMM_ID_BYTES = 128
REC_HEAD = 144   # offsetof(zkb_result, parts)


def _preload_nccl():
    """The C library opens NCCL by its SONAME at the first multi-GPU call.  In a Python process that may import torch LATER,
    the system libnccl.so.2 must not be the copy that gets mapped: torch's libtorch_cuda.so needs the (newer) NCCL its wheel
    bundles and the loader would hand it the one already mapped (ImportError: undefined symbol).  So the bundled copy, if
    there is one, is mapped first; non-Python callers just get the system library."""
    try:
        import glob
        import importlib.util
        import os
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
            for f in sorted(glob.glob(os.path.join(d, "lib", "libnccl.so*"))):
                C.CDLL(f, mode=C.RTLD_GLOBAL)
                return
    except Exception:   # noqa: BLE001  (no bundled copy / not loadable: the SONAME lookup of the C library decides)
        pass


def _bind(L):
    if getattr(L, "_multi_bound", False):
        return L
    _preload_nccl()
    vp, sz = C.c_void_p, C.c_size_t
    L.zkb_plan_shards.argtypes = [vp, sz, sz, C.c_int, C.POINTER(sz)]
    L.zkb_multi_create.argtypes = [C.POINTER(_Options), C.POINTER(C.c_int32), sz, C.POINTER(vp)]
    L.zkb_multi_destroy.argtypes = [vp]
    L.zkb_multi_destroy.restype = None
    L.zkb_multi_devices.argtypes = [vp]
    L.zkb_multi_devices.restype = sz
    L.zkb_multi_engine.argtypes = [vp, sz]
    L.zkb_multi_engine.restype = vp
    L.zkb_multi_host_register.argtypes = [vp, vp, sz]
    L.zkb_multi_host_unregister.argtypes = [vp, vp]
    L.zkb_multi_regex_set_create.argtypes = [vp, C.POINTER(_DfaView), sz, sz, C.c_int, C.c_int, C.POINTER(vp)]
    L.zkb_multi_regex_destroy.argtypes = [vp]
    L.zkb_multi_regex_destroy.restype = None
    L.zkb_multi_verify_batch.argtypes = [vp, vp, sz, vp, vp, vp]
    L.zkb_multi_batch_prepare.argtypes = [vp, vp, sz, vp, vp, C.c_int, C.POINTER(vp)]
    L.zkb_multi_batch_run.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
    L.zkb_multi_batch_fetch.argtypes = [vp, vp]
    L.zkb_multi_batch_bounds.argtypes = [vp, C.POINTER(sz), sz]
    L.zkb_multi_batch_gathered.argtypes = [vp, sz, vp, C.POINTER(sz)]
    L.zkb_multi_batch_destroy.argtypes = [vp]
    L.zkb_multi_batch_destroy.restype = None
    L.zkb_comm_unique_id.argtypes = [vp]
    L.zkb_comm_create.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.zkb_comm_destroy.argtypes = [vp]
    L.zkb_comm_destroy.restype = None
    L.zkb_comm_allgather_records.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(sz), C.POINTER(sz)]
    L.zkb_comm_run_allgather.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(sz), C.POINTER(sz)]
    L.zkb_comm_rank_records.argtypes = [vp, C.POINTER(C.c_uint64), sz]
    L._multi_bound = True
    return L


def plan_shards(views: EmailViews, n_shards: int, resident: bool = False) -> List[int]:
    """Contiguous cost-balanced shards: shard k = emails [b[k], b[k+1]).  Host-only (zkb_plan_shards)."""
    L = _bind(load_library())
    b = (C.c_size_t * (n_shards + 1))()
    _check(L.zkb_plan_shards(views.ptr, views.n, n_shards, 1 if resident else 0, b), "zkb_plan_shards")
    return [int(x) for x in b]


def records_to_results(recs: np.ndarray, n_parts: int) -> np.ndarray:
    """(n, 144 + 16 * P) device-built record bytes -> RESULT_DTYPE records (the head is the head of zkb_result)."""
    n = recs.shape[0]
    out = np.zeros(n, dtype=RESULT_DTYPE)
    raw = out.view(np.uint8).reshape(n, RESULT_DTYPE.itemsize)
    raw[:, :REC_HEAD] = recs[:, :REC_HEAD]
    raw[:, REC_HEAD:REC_HEAD + 16 * n_parts] = recs[:, REC_HEAD:REC_HEAD + 16 * n_parts]
    return out


class MultiRegexSet:
    def __init__(self, multi: "MultiEngine", info: RegexInfo):
        self.multi = multi
        # reuse RegexSet's marshalling of the parts without creating a device set: build the views here
        hp, bp = info.header_parts or [], info.body_parts or []
        self.parts = list(hp) + list(bp)
        arr = (_DfaView * max(1, len(self.parts)))()
        self._keep = []
        for i, p in enumerate(self.parts):
            f, b = bytes(p.verify_re.fwd), bytes(p.verify_re.bwd)
            self._keep += [f, b]
            arr[i].fwd = C.cast(C.c_char_p(f), C.c_void_p).value
            arr[i].fwd_len = len(f)
            arr[i].bwd = C.cast(C.c_char_p(b), C.c_void_p).value
            arr[i].bwd_len = len(b)
        self.handle = C.c_void_p()
        self.n_active = (len(hp) if info.header_parts is not None else 0) + (len(bp) if info.body_parts is not None else 0)
        _check(multi.lib.zkb_multi_regex_set_create(multi.handle, arr, len(hp), len(bp), 1 if info.header_parts is not None else 0,
                                                    1 if info.body_parts is not None else 0, C.byref(self.handle)),
               "zkb_multi_regex_set_create")

    def close(self):
        if self.handle:
            self.multi.lib.zkb_multi_regex_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiBatch:
    def __init__(self, multi: "MultiEngine", views: EmailViews, regex: Optional[MultiRegexSet], raw: bool):
        self.multi, self.views, self.regex = multi, views, regex
        self.handle = C.c_void_p()
        _check(multi.lib.zkb_multi_batch_prepare(multi.handle, views.ptr, views.n, regex.handle if regex else None, None,
                                                 1 if raw else 0, C.byref(self.handle)), "zkb_multi_batch_prepare")

    def bounds(self) -> List[int]:
        d = self.multi.n_devices
        b = (C.c_size_t * (d + 1))()
        _check(self.multi.lib.zkb_multi_batch_bounds(self.handle, b, d + 1), "zkb_multi_batch_bounds")
        return [int(x) for x in b]

    def run(self, gather: bool = True) -> float:
        """All shards' kernels (+ the all-gather of the result records); returns the slowest device's time in ms."""
        ms = C.c_float()
        _check(self.multi.lib.zkb_multi_batch_run(self.handle, 1 if gather else 0, C.byref(ms)), "zkb_multi_batch_run")
        return float(ms.value)

    def fetch(self) -> np.ndarray:
        out = np.zeros(self.views.n, dtype=RESULT_DTYPE)
        _check(self.multi.lib.zkb_multi_batch_fetch(self.handle, out.ctypes.data), "zkb_multi_batch_fetch")
        return out

    def gathered(self, device_index: int) -> np.ndarray:
        """The whole batch's records as device `device_index` holds them after run(gather=True)."""
        rb = C.c_size_t()
        _check(self.multi.lib.zkb_multi_batch_gathered(self.handle, device_index, None, C.byref(rb)), "zkb_multi_batch_gathered")
        recs = np.zeros((self.views.n, rb.value), dtype=np.uint8)
        _check(self.multi.lib.zkb_multi_batch_gathered(self.handle, device_index, recs.ctypes.data, C.byref(rb)), "zkb_multi_batch_gathered")
        return records_to_results(recs, self.regex.n_active if self.regex else 0)

    def close(self):
        if self.handle:
            self.multi.lib.zkb_multi_batch_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiEngine:
    """One process, several devices (zkb_multi): one engine, stream set and host-thread group per device."""

    def __init__(self, devices: Optional[Sequence[int]] = None, n_devices: int = 0, host_threads: int = 0, now_unix: int = 0,
                 chunk_emails: int = 0, flags: int = 0):
        self.lib = _bind(load_library())
        devs = list(devices) if devices is not None else list(range(n_devices or 1))
        arr = (C.c_int32 * len(devs))(*devs)
        opt = _Options(0, host_threads, now_unix, chunk_emails, flags, 0)
        self.handle = C.c_void_p()
        self.now_unix = now_unix
        _check(self.lib.zkb_multi_create(C.byref(opt), arr, len(devs), C.byref(self.handle)), "zkb_multi_create")
        self.n_devices = len(devs)

    def close(self):
        if self.handle:
            self.lib.zkb_multi_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def register_host(self, array: np.ndarray):
        _check(self.lib.zkb_multi_host_register(self.handle, array.ctypes.data, array.nbytes), "zkb_multi_host_register")

    def unregister_host(self, array: np.ndarray):
        _check(self.lib.zkb_multi_host_unregister(self.handle, array.ctypes.data), "zkb_multi_host_unregister")

    def verify_views(self, views: EmailViews, regex: Optional[MultiRegexSet] = None) -> np.ndarray:
        out = np.zeros(views.n, dtype=RESULT_DTYPE)
        _check(self.lib.zkb_multi_verify_batch(self.handle, views.ptr, views.n, regex.handle if regex else None, None,
                                               out.ctypes.data), "zkb_multi_verify_batch")
        return out

    def verify_batch(self, emails: Sequence[Email]) -> np.ndarray:
        return self.verify_views(EmailViews.from_emails(emails))

    def verify_with_regex_batch(self, emails: Sequence[Email], info: RegexInfo) -> np.ndarray:
        rs = MultiRegexSet(self, info)
        try:
            return self.verify_views(EmailViews.from_emails(emails), rs)
        finally:
            rs.close()

    def prepare(self, views: EmailViews, regex: Optional[MultiRegexSet] = None, raw: bool = False) -> MultiBatch:
        return MultiBatch(self, views, regex, raw)


class Comm:
    """One process per device: NCCL communicator owned by the library (zkb_comm).  `broadcast` carries the 128-byte id
    from rank 0 to the other ranks (default: torch.distributed.broadcast_object_list on the default process group)."""

    def __init__(self, engine: Engine, rank: int, world: int, broadcast=None):
        self.lib = _bind(load_library())
        self.engine, self.rank, self.world = engine, rank, world
        ident = (C.c_uint8 * COMM_ID_BYTES)()
        if rank == 0:
            _check(self.lib.zkb_comm_unique_id(ident), "zkb_comm_unique_id")
        if world > 1:
            if broadcast is None:
                import torch.distributed as dist

                def broadcast(b):
                    box = [b]
                    dist.broadcast_object_list(box, src=0)
                    return box[0]
            got = broadcast(bytes(ident) if rank == 0 else None)
            ident = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(got)
        self.handle = C.c_void_p()
        _check(self.lib.zkb_comm_create(engine.handle, ident, rank, world, C.byref(self.handle)), "zkb_comm_create")

    def allgather_records(self, batch: PreparedBatch):
        """Enqueues the all-gather of the batch's records behind its last run; returns (device pointer, slot_records,
        rec_bytes): rank r's records start at r * slot_records."""
        p, slot, rb = C.c_void_p(), C.c_size_t(), C.c_size_t()
        _check(self.lib.zkb_comm_allgather_records(self.handle, batch.handle, C.byref(p), C.byref(slot), C.byref(rb)),
               "zkb_comm_allgather_records")
        return p.value, slot.value, rb.value

    def run_allgather(self, batch: PreparedBatch):
        """batch.run_async() and the exchange of the records in one call: chunk k's records travel while chunk k + 1 is
        computed.  Same return value as allgather_records; every rank must call it."""
        p, slot, rb = C.c_void_p(), C.c_size_t(), C.c_size_t()
        _check(self.lib.zkb_comm_run_allgather(self.handle, batch.handle, C.byref(p), C.byref(slot), C.byref(rb)),
               "zkb_comm_run_allgather")
        return p.value, slot.value, rb.value

    def rank_records(self) -> List[int]:
        a = (C.c_uint64 * self.world)()
        _check(self.lib.zkb_comm_rank_records(self.handle, a, self.world), "zkb_comm_rank_records")
        return [int(x) for x in a]

    def close(self):
        if self.handle:
            self.lib.zkb_comm_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
