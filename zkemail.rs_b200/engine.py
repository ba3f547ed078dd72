"""Host-side mirror of the reference's operator interface for the hot path, on top of the C ABI of
``libzkemail_b200.so`` (include/zkemail_b200.h).

Reference surface mirrored (same names, argument meaning and error behaviour):

* ``verify_email(&Email) -> EmailVerifierOutput``                        core/src/circuits.rs:9-29
* ``verify_email_with_regex(&EmailWithRegex) -> EmailWithRegexVerifierOutput``   circuits.rs:31-68
* every ``unwrap``/``assert!`` panic site becomes :class:`VerificationPanic` carrying the status
  code; the batch entry points return per-email records instead of raising.

New (the point of the engine): ``Engine.verify_batch`` / ``Engine.verify_with_regex_batch``.

The arithmetic runs ONLY in the CUDA library.  There is no CPU fallback: constructing an
:class:`Engine` without a usable B200 raises :class:`EngineUnavailable`.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .structs import (
    DFA, CompiledRegex, Email, EmailVerifierOutput, EmailWithRegex, EmailWithRegexVerifierOutput,
    RegexInfo, RegexPattern,
)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzkemail_b200.so")
ZKB_MAX_PARTS = 16

STATUS_NAMES = {
    0: "OK", 1: "MAIL_PARSE", 2: "KEY", 3: "DKIM_FAIL", 4: "NULL_EXTERNAL", 5: "CANONICALIZE",
    6: "REGEX_HEADER", 7: "REGEX_BODY", 8: "BAD_DFA", 9: "UNSUPPORTED",
}
# the reference's panic message / site for each status (for error text only)
_PANIC_SITE = {
    1: "core/src/email.rs:26 parse_mail(..).unwrap()",
    2: "core/src/email.rs:28-29 DkimPublicKey::try_from_bytes(..).unwrap()",
    3: "core/src/circuits.rs:13 assert!(verified)",
    4: "core/src/circuits.rs:24 Value cannot be null",
    5: "core/src/circuits.rs:35 canonicalize_signed_email(..).unwrap()",
    6: "core/src/circuits.rs:45 assert!(verified) [header regex]",
    7: "core/src/circuits.rs:54 assert!(verified) [body regex]",
    8: "core/src/regex.rs:32-33 DFA::from_bytes(..).unwrap()",
    9: "unsupported algorithm (ed25519 key / rsa-sha1): declined by the engine",
}


class EngineUnavailable(RuntimeError):
    """The CUDA library or a usable device is missing.  There is no CPU fallback."""


class VerificationPanic(AssertionError):
    """Raised by the single-email wrappers where the reference panics."""

    def __init__(self, status: int, detail: int = 0):
        self.status, self.detail = status, detail
        super().__init__(f"{STATUS_NAMES.get(status, status)} (dkim detail {detail}): "
                         f"{_PANIC_SITE.get(status, '')}")


class RegexError(ValueError):
    pass


# ------------------------------------------------------------------ ctypes mirrors of the header
class _EmailView(C.Structure):
    _fields_ = [("from_domain", C.c_void_p), ("from_domain_len", C.c_size_t),
                ("raw_email", C.c_void_p), ("raw_email_len", C.c_size_t),
                ("key", C.c_void_p), ("key_len", C.c_size_t),
                ("key_type", C.c_void_p), ("key_type_len", C.c_size_t)]


class _Part(C.Structure):
    _fields_ = [("match_count", C.c_uint32), ("start", C.c_uint32), ("end", C.c_uint32),
                ("captures_ok", C.c_uint32)]


class Result(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("dkim_detail", C.c_int32),
        ("body_hash", C.c_uint8 * 32), ("header_hash", C.c_uint8 * 32),
        ("from_domain_hash", C.c_uint8 * 32), ("public_key_hash", C.c_uint8 * 32),
        ("bh_ok", C.c_uint8), ("rsa_ok", C.c_uint8), ("pad", C.c_uint8 * 2),
        ("n_parts", C.c_uint32), ("parts", _Part * ZKB_MAX_PARTS),
    ]

    def as_dict(self) -> dict:
        return {
            "status": self.status, "dkim_detail": self.dkim_detail,
            "body_hash": bytes(self.body_hash), "header_hash": bytes(self.header_hash),
            "from_domain_hash": bytes(self.from_domain_hash),
            "public_key_hash": bytes(self.public_key_hash),
            "bh_ok": int(self.bh_ok), "rsa_ok": int(self.rsa_ok),
            "parts": [(p.match_count, p.start, p.end, p.captures_ok)
                      for p in list(self.parts)[: self.n_parts]],
        }


RESULT_DTYPE = np.dtype([
    ("status", "<i4"), ("dkim_detail", "<i4"), ("body_hash", "u1", 32), ("header_hash", "u1", 32),
    ("from_domain_hash", "u1", 32), ("public_key_hash", "u1", 32), ("bh_ok", "u1"), ("rsa_ok", "u1"),
    ("pad", "u1", 2), ("n_parts", "<u4"), ("parts", "<u4", (ZKB_MAX_PARTS, 4)),
])
assert RESULT_DTYPE.itemsize == C.sizeof(Result)


class _DfaView(C.Structure):
    _fields_ = [("fwd", C.c_void_p), ("fwd_len", C.c_size_t), ("bwd", C.c_void_p), ("bwd_len", C.c_size_t)]


class _Capture(C.Structure):
    _fields_ = [("part", C.c_uint32), ("s", C.c_void_p), ("len", C.c_size_t)]


class _EmailCaptures(C.Structure):
    _fields_ = [("caps", C.POINTER(_Capture)), ("n_caps", C.c_size_t)]


class _Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("host_threads", C.c_int32), ("now_unix", C.c_int64),
                ("chunk_emails", C.c_uint64), ("flags", C.c_uint32), ("rsa_lanes", C.c_uint32)]


class BatchStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "n_emails", "n_candidates", "n_sha_messages", "sha_blocks", "sha_bytes", "rsa_items_1024",
        "rsa_items_2048", "rsa_items_other", "rsa_macs", "dfa_items", "dfa_bytes", "arena_bytes",
        "h2d_bytes", "d2h_bytes", "kernel_launches")]

    def as_dict(self) -> dict:
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


EXPORTED_SYMBOLS = (
    "zkb_abi_version", "zkb_strerror", "zkb_engine_create", "zkb_engine_destroy",
    "zkb_regex_set_create", "zkb_regex_set_destroy", "zkb_verify_batch", "zkb_verify_one",
    "zkb_batch_prepare", "zkb_batch_run", "zkb_batch_run_async", "zkb_batch_fetch",
    "zkb_batch_destroy", "zkb_batch_get_stats", "zkb_batch_last_timing", "zkb_engine_stream",
    "zkb_regex_compile", "zkb_free", "zkb_sha256_batch", "zkb_rsa_verify_batch",
    "zkb_dfa_scan_batch", "zkb_int_pipe_peaks", "zkb_host_canonicalize", "zkb_batch_device_flags",
    "zkb_host_register", "zkb_host_unregister", "zkb_engine_last_batch_bytes",
    "zkb_abi_encode_batch", "zkb_abi_decode", "zkb_host_dkim_signatures",
    "zkb_regex_automata_to_zdf", "zkb_engine_set_flags", "zkb_batch_prepare_raw", "zkb_batch_last_timing_ex",
    "zkb_plan_shards", "zkb_multi_create", "zkb_multi_destroy", "zkb_multi_devices", "zkb_multi_engine", "zkb_multi_host_register",
    "zkb_multi_host_unregister", "zkb_multi_regex_set_create", "zkb_multi_regex_destroy", "zkb_multi_verify_batch",
    "zkb_multi_batch_prepare", "zkb_multi_batch_run", "zkb_multi_batch_fetch", "zkb_multi_batch_bounds", "zkb_multi_batch_gathered",
    "zkb_multi_batch_destroy", "zkb_comm_unique_id", "zkb_comm_create", "zkb_comm_destroy", "zkb_comm_allgather_records", "zkb_comm_run_allgather",
    "zkb_comm_rank_records",
)

# zkb_options.flags / zkb_engine_set_flags (include/zkemail_b200.h)
OPT_NO_DIRECT, OPT_NO_DEVICE_FRONTEND, OPT_NO_STAGED_FRONTEND, OPT_NO_OVERLAP, OPT_PROFILE, OPT_NO_SQR = 1, 2, 4, 8, 16, 32

_lib = None


def build_library(force: bool = False) -> str:
    """Compiles libzkemail_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src = os.path.join(_HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-C", src, "-s", "clean"])
    subprocess.check_call(["make", "-C", src, "-s", "-j", str(min(8, os.cpu_count() or 1))])
    return LIB_PATH


def load_library():
    """Loads the CUDA library; raises EngineUnavailable when it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineUnavailable(f"{LIB_PATH} is not built (run __graft_entry__.build()); "
                                "zkemail_b200 has no CPU fallback")
    try:
        L = C.CDLL(LIB_PATH)
    except OSError as e:  # e.g. libcudart missing
        raise EngineUnavailable(f"cannot load {LIB_PATH}: {e}") from e
    vp, sz = C.c_void_p, C.c_size_t
    L.zkb_abi_version.restype = C.c_int
    L.zkb_strerror.restype = C.c_char_p
    L.zkb_strerror.argtypes = [C.c_int]
    L.zkb_engine_create.argtypes = [C.POINTER(_Options), C.POINTER(vp)]
    L.zkb_engine_destroy.argtypes = [vp]
    L.zkb_engine_destroy.restype = None
    L.zkb_engine_set_flags.argtypes = [vp, C.c_uint32]
    L.zkb_regex_set_create.argtypes = [vp, C.POINTER(_DfaView), sz, sz, C.c_int, C.c_int, C.POINTER(vp)]
    L.zkb_regex_set_destroy.argtypes = [vp]
    L.zkb_regex_set_destroy.restype = None
    L.zkb_verify_batch.argtypes = [vp, vp, sz, vp, vp, vp]
    L.zkb_verify_one.argtypes = [vp, vp, vp, vp, vp]
    L.zkb_batch_prepare.argtypes = [vp, vp, sz, vp, vp, C.POINTER(vp)]
    L.zkb_batch_prepare_raw.argtypes = [vp, vp, sz, vp, vp, C.POINTER(vp)]
    L.zkb_batch_last_timing_ex.argtypes = [vp, C.POINTER(C.c_float), sz]
    L.zkb_batch_run.argtypes = [vp]
    L.zkb_batch_run_async.argtypes = [vp]
    L.zkb_batch_fetch.argtypes = [vp, vp]
    L.zkb_batch_destroy.argtypes = [vp]
    L.zkb_batch_destroy.restype = None
    L.zkb_batch_get_stats.argtypes = [vp, C.POINTER(BatchStats)]
    L.zkb_batch_last_timing.argtypes = [vp, C.POINTER(C.c_float * 5)]
    L.zkb_batch_device_flags.argtypes = [vp, sz, C.POINTER(vp), C.POINTER(sz), C.POINTER(sz)]
    L.zkb_engine_last_batch_bytes.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.zkb_host_register.argtypes = [vp, vp, sz]
    L.zkb_host_unregister.argtypes = [vp, vp]
    L.zkb_engine_stream.argtypes = [vp]
    L.zkb_engine_stream.restype = vp
    L.zkb_regex_compile.argtypes = [C.c_char_p, sz, C.POINTER(vp), C.POINTER(sz), C.POINTER(vp),
                                    C.POINTER(sz), C.c_char_p, sz]
    L.zkb_free.argtypes = [vp]
    L.zkb_free.restype = None
    L.zkb_sha256_batch.argtypes = [vp, vp, sz, vp, vp, sz, vp]
    L.zkb_rsa_verify_batch.argtypes = [vp, vp, vp, vp, vp, vp, sz, vp]
    L.zkb_dfa_scan_batch.argtypes = [vp, C.POINTER(_DfaView), vp, sz, vp, vp, sz, C.c_int, vp]
    L.zkb_int_pipe_peaks.argtypes = [vp, C.POINTER(C.c_double * 8)]
    L.zkb_host_canonicalize.argtypes = [C.c_char_p, sz, C.c_int64, C.POINTER(vp), C.POINTER(sz), C.POINTER(vp),
                                        C.POINTER(sz), C.POINTER(C.c_int)]
    _lib = L
    return L


def _check(rc: int, what: str):
    if rc != 0:
        msg = load_library().zkb_strerror(rc).decode()
        if rc == 2:
            raise EngineUnavailable(f"{what}: {msg}")
        if rc == 5:
            raise RegexError(f"{what}: {msg}")
        raise RuntimeError(f"{what}: {msg} (code {rc})")


# ------------------------------------------------------------------ regex compiler front end
def compile_regex(pattern: str) -> DFA:
    """pattern -> DFA{fwd,bwd} (ZDF1 tables); stands in for helpers/src/regex.rs:7-14 create_dfa."""
    L = load_library()
    p = pattern.encode("utf-8")
    f, b = C.c_void_p(), C.c_void_p()
    fl, bl = C.c_size_t(), C.c_size_t()
    err = C.create_string_buffer(256)
    rc = L.zkb_regex_compile(p, len(p), C.byref(f), C.byref(fl), C.byref(b), C.byref(bl), err, 256)
    if rc:
        raise RegexError(f"{pattern!r}: {err.value.decode(errors='replace')}")
    try:
        return DFA(C.string_at(f, fl.value), C.string_at(b, bl.value))
    finally:
        L.zkb_free(f)
        L.zkb_free(b)


def _py_pattern(pattern: str) -> bytes:
    """Translates the few Rust-only spellings to Python `re` (host-side capture resolution only)."""
    posix = {"alnum": "0-9A-Za-z", "alpha": "A-Za-z", "digit": "0-9", "lower": "a-z", "upper": "A-Z",
             "space": r"\t\n\v\f\r ", "xdigit": "0-9A-Fa-f", "word": r"0-9A-Za-z_", "blank": r"\t ",
             "punct": r"!-/:-@\[-`{-~"}
    p = pattern
    for k, v in posix.items():
        p = p.replace(f"[:{k}:]", v)
    p = p.replace(r"\z", r"\Z").replace("(?<", "(?P<") if "(?<=" not in p and "(?<!" not in p else p

    # inline flag groups: Python's bytes patterns are ASCII-only and have no `u` flag to set or clear
    def _flags(m):
        on, off, tail = m.group(1), (m.group(2) or "")[1:], m.group(3)
        on, off = on.replace("u", ""), off.replace("u", "")
        if "U" in on or "U" in off:
            raise RegexError("the swap-greed flag U has no Python equivalent (capture resolution)")
        body = on + ("-" + off if off else "")
        if not body:
            return "(?:" if tail == ":" else ""
        return "(?" + body + tail
    p = re.sub(r"\(\?([a-zA-Z]*)(-[a-zA-Z]*)?([:)])", _flags, p)
    return p.encode("utf-8")


def regex_automata_to_zdf(wire: bytes, reverse: bool) -> bytes:
    """regex-automata dense-DFA bytes (helpers/src/regex.rs:7-14) -> the engine's ZDF1 table; raises RegexError
    where dense::DFA::from_bytes would fail (core/src/regex.rs:32-33).  The engine does this itself on load."""
    L = load_library()
    L.zkb_regex_automata_to_zdf.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    out, n = C.c_void_p(), C.c_size_t()
    if L.zkb_regex_automata_to_zdf(wire, len(wire), 1 if reverse else 0, C.byref(out), C.byref(n)) != 0:
        raise RegexError("not a valid regex-automata dense DFA")
    try:
        return C.string_at(out.value, n.value)
    finally:
        L.zkb_free(out)


def compile_regex_parts(parts: Sequence[RegexPattern], haystack: bytes) -> List[CompiledRegex]:
    """Mirror of helpers/src/regex.rs:16-51 `compile_regex_parts`: each pattern must match the
    input exactly once; `capture_indices` are resolved to capture STRINGS on the host (the
    reference uses a meta::Regex for that step; here Python `re`, leftmost-first like Rust)."""
    out = []
    for part in parts:
        dfa = compile_regex(part.pattern)
        rx = re.compile(_py_pattern(part.pattern))
        ms = [m for m in rx.finditer(haystack)]
        if len(ms) != 1:
            raise RegexError(f"Input doesn't match regex pattern exactly once: {part.pattern!r} "
                             f"({len(ms)} matches)")
        caps: List[str] = []
        if part.capture_indices:
            m = ms[0]
            for idx in part.capture_indices:
                if idx < 1 or idx > (rx.groups or 0) or m.group(idx) is None:
                    raise RegexError(f"Capture group {idx} not found in match for {part.pattern!r}")
                caps.append(m.group(idx).decode("utf-8", errors="replace"))
        out.append(CompiledRegex(verify_re=dfa, captures=caps))
    return out


# ------------------------------------------------------------------ batches of borrowed views
class EmailViews:
    """`&[Email]` as the C ABI sees it: an array of zkb_email_view over buffers kept alive here."""

    def __init__(self, arr, n: int, keep):
        self.arr, self.n, self._keep = arr, n, keep

    @property
    def ptr(self) -> int:
        return C.addressof(self.arr) if not isinstance(self.arr, np.ndarray) else self.arr.ctypes.data

    @staticmethod
    def from_emails(emails: Sequence[Email]) -> "EmailViews":
        n = len(emails)
        arr = (_EmailView * max(1, n))()
        keep = []
        for i, e in enumerate(emails):
            dom = e.from_domain.encode("utf-8")
            raw = bytes(e.raw_email)
            key = bytes(e.public_key.key)
            kt = e.public_key.key_type.encode("utf-8")
            keep += [dom, raw, key, kt]
            v = arr[i]
            v.from_domain = C.cast(C.c_char_p(dom), C.c_void_p).value
            v.from_domain_len = len(dom)
            v.raw_email = C.cast(C.c_char_p(raw), C.c_void_p).value
            v.raw_email_len = len(raw)
            v.key = C.cast(C.c_char_p(key), C.c_void_p).value
            v.key_len = len(key)
            v.key_type = C.cast(C.c_char_p(kt), C.c_void_p).value
            v.key_type_len = len(kt)
        return EmailViews(arr, n, keep)

    @staticmethod
    def from_arrays(views: np.ndarray, keep) -> "EmailViews":
        """views: (n, 8) uint64 array already laid out as zkb_email_view records."""
        assert views.dtype == np.uint64 and views.ndim == 2 and views.shape[1] == 8 and views.flags.c_contiguous
        return EmailViews(views, views.shape[0], keep)


class RegexSet:
    """Device-resident DFAs of one RegexInfo (core/src/structs.rs:32-35), shared by a batch."""

    def __init__(self, engine: "Engine", info: RegexInfo):
        self.engine = engine
        hp = info.header_parts or []
        bp = info.body_parts or []
        self.header_present = info.header_parts is not None
        self.body_present = info.body_parts is not None
        self.n_header, self.n_body = len(hp), len(bp)
        self.parts = list(hp) + list(bp)
        views = (_DfaView * max(1, len(self.parts)))()
        self._keep = []
        for i, p in enumerate(self.parts):
            f, b = bytes(p.verify_re.fwd), bytes(p.verify_re.bwd)
            self._keep += [f, b]
            views[i].fwd = C.cast(C.c_char_p(f), C.c_void_p).value
            views[i].fwd_len = len(f)
            views[i].bwd = C.cast(C.c_char_p(b), C.c_void_p).value
            views[i].bwd_len = len(b)
        self.handle = C.c_void_p()
        rc = engine.lib.zkb_regex_set_create(engine.handle, views, self.n_header, self.n_body,
                                             1 if self.header_present else 0,
                                             1 if self.body_present else 0, C.byref(self.handle))
        if rc == 5:
            raise VerificationPanic(8)  # DFA::from_bytes(..).unwrap()
        _check(rc, "zkb_regex_set_create")

    def captures_for(self, n: int):
        """zkb_email_captures array giving every email of a batch this set's capture strings."""
        caps = []
        for pi, p in enumerate(self.parts):
            for s in (p.captures or []):
                caps.append((pi, s.encode("utf-8")))
        if not caps:
            return None, None
        arr = (_Capture * len(caps))()
        keep = [arr]
        for i, (pi, s) in enumerate(caps):
            keep.append(s)
            arr[i].part = pi
            arr[i].s = C.cast(C.c_char_p(s), C.c_void_p).value
            arr[i].len = len(s)
        ec = np.empty((max(1, n), 2), dtype=np.uint64)  # zkb_email_captures records
        ec[:, 0] = C.addressof(arr)
        ec[:, 1] = len(caps)
        return ec, keep

    def close(self):
        if self.handle:
            self.engine.lib.zkb_regex_set_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PreparedBatch:
    """A batch packed and resident in HBM (zkb_batch): run() launches only the kernels."""

    def __init__(self, engine: "Engine", views: EmailViews, regex: Optional[RegexSet], caps, raw: bool = False):
        self.engine, self.views, self.regex, self._caps, self.raw = engine, views, regex, caps, raw
        self.handle = C.c_void_p()
        cap_ptr = caps[0].ctypes.data if caps and caps[0] is not None else None
        fn = engine.lib.zkb_batch_prepare_raw if raw else engine.lib.zkb_batch_prepare
        _check(fn(engine.handle, views.ptr, views.n, regex.handle if regex else None, cap_ptr,
                  C.byref(self.handle)), "zkb_batch_prepare_raw" if raw else "zkb_batch_prepare")

    def run(self):
        _check(self.engine.lib.zkb_batch_run(self.handle), "zkb_batch_run")

    def run_async(self):
        _check(self.engine.lib.zkb_batch_run_async(self.handle), "zkb_batch_run_async")

    def timing_ms(self) -> dict:
        t = (C.c_float * 8)()
        _check(self.engine.lib.zkb_batch_last_timing_ex(self.handle, t, 8), "zkb_batch_last_timing_ex")
        return {"sha256": t[0], "rsa": t[1], "dfa": t[2], "bh_check": t[3], "total": t[4], "front_end_canon": t[5], "records": t[6]}

    def stats(self) -> dict:
        s = BatchStats()
        _check(self.engine.lib.zkb_batch_get_stats(self.handle, C.byref(s)), "zkb_batch_get_stats")
        return s.as_dict()

    def device_flags(self):
        """[(device pointer, count)] of the per-candidate verdict words of every resident chunk."""
        nch = C.c_size_t()
        _check(self.engine.lib.zkb_batch_device_flags(self.handle, 1 << 60, None, None, C.byref(nch)), "zkb_batch_device_flags")
        out = []
        for i in range(nch.value):
            p, n = C.c_void_p(), C.c_size_t()
            _check(self.engine.lib.zkb_batch_device_flags(self.handle, i, C.byref(p), C.byref(n), None), "zkb_batch_device_flags")
            out.append((p.value, n.value))
        return out

    def fetch(self) -> np.ndarray:
        out = np.zeros(self.views.n, dtype=RESULT_DTYPE)
        _check(self.engine.lib.zkb_batch_fetch(self.handle, out.ctypes.data), "zkb_batch_fetch")
        return out

    def close(self):
        if self.handle:
            self.engine.lib.zkb_batch_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One engine per process and device (zkb_engine)."""

    def __init__(self, device: int = 0, host_threads: int = 0, now_unix: int = 0,
                 chunk_emails: int = 0, rsa_lanes: int = 0, flags: int = 0):
        self.lib = load_library()
        self.now_unix = now_unix
        self.flags = flags
        opt = _Options(device, host_threads, now_unix, chunk_emails, flags, rsa_lanes)
        self.handle = C.c_void_p()
        _check(self.lib.zkb_engine_create(C.byref(opt), C.byref(self.handle)), "zkb_engine_create")

    def close(self):
        if self.handle:
            self.lib.zkb_engine_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_flags(self, flags: int):
        """Path switches (OPT_* bits); every path gives identical results."""
        _check(self.lib.zkb_engine_set_flags(self.handle, flags), "zkb_engine_set_flags")
        self.flags = flags

    def last_batch_bytes(self) -> dict:
        h, d, f = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(self.lib.zkb_engine_last_batch_bytes(self.handle, C.byref(h), C.byref(d), C.byref(f)), "zkb_engine_last_batch_bytes")
        return {"h2d_bytes": h.value, "d2h_bytes": d.value, "host_front_end_emails": f.value}

    def register_host(self, array: np.ndarray):
        """Page-locks caller memory holding raw messages (zero-copy input path, zkb_host_register)."""
        _check(self.lib.zkb_host_register(self.handle, array.ctypes.data, array.nbytes), "zkb_host_register")

    def unregister_host(self, array: np.ndarray):
        _check(self.lib.zkb_host_unregister(self.handle, array.ctypes.data), "zkb_host_unregister")

    # ---- batch entry points -------------------------------------------------------------
    def verify_views(self, views: EmailViews, regex: Optional[RegexSet] = None, with_captures: bool = True,
                     out: Optional[np.ndarray] = None) -> np.ndarray:
        """zkb_verify_batch.  `out`: caller-owned result array to reuse (RESULT_DTYPE, views.n records)."""
        if out is None:
            out = np.zeros(views.n, dtype=RESULT_DTYPE)
        assert out.dtype == RESULT_DTYPE and len(out) == views.n and out.flags.c_contiguous
        caps = regex.captures_for(views.n) if (regex and with_captures) else (None, None)
        cap_ptr = caps[0].ctypes.data if caps[0] is not None else None
        _check(self.lib.zkb_verify_batch(self.handle, views.ptr, views.n, regex.handle if regex else None,
                                         cap_ptr, out.ctypes.data), "zkb_verify_batch")
        return out

    def verify_batch(self, emails: Sequence[Email]) -> np.ndarray:
        """verify_email over a batch: one record per email (status 0 = the reference returns)."""
        return self.verify_views(EmailViews.from_emails(emails))

    def verify_with_regex_batch(self, emails: Sequence[Email], regex_info: RegexInfo) -> np.ndarray:
        """verify_email_with_regex over a batch sharing one RegexInfo."""
        rs = RegexSet(self, regex_info)
        try:
            return self.verify_views(EmailViews.from_emails(emails), rs)
        finally:
            rs.close()

    def prepare(self, views: EmailViews, regex: Optional[RegexSet] = None, with_captures: bool = True,
                raw: bool = False) -> PreparedBatch:
        """Resident batch.  raw=False: canonical bytes packed in HBM, run() = hashing + RSA + DFA scans.  raw=True: the raw
        messages are resident and run() starts at the device front end (zkb_batch_prepare_raw)."""
        caps = regex.captures_for(views.n) if (regex and with_captures) else (None, None)
        return PreparedBatch(self, views, regex, caps, raw)

    # ---- the reference's single-email call shape ------------------------------------------
    def verify_email(self, email: Email) -> EmailVerifierOutput:
        r = self.verify_batch([email])[0]
        if r["status"] != 0:
            raise VerificationPanic(int(r["status"]), int(r["dkim_detail"]))
        return _email_output(email, r)

    def verify_email_with_regex(self, inp: EmailWithRegex) -> EmailWithRegexVerifierOutput:
        r = self.verify_with_regex_batch([inp.email], inp.regex_info)[0]
        if r["status"] != 0:
            raise VerificationPanic(int(r["status"]), int(r["dkim_detail"]))
        matches: List[str] = []
        for plist in (inp.regex_info.header_parts, inp.regex_info.body_parts):
            for p in (plist or []):
                matches.extend(p.captures or [])
        return EmailWithRegexVerifierOutput(_email_output(inp.email, r), matches)

    # ---- kernel-level entry points ----------------------------------------------------------
    def sha256_batch(self, messages: Sequence[bytes]) -> List[bytes]:
        n = len(messages)
        data = b"".join(messages)
        lens = np.array([len(m) for m in messages], dtype=np.uint32)
        offs = np.zeros(n, dtype=np.uint64)
        if n:
            offs[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
        out = np.zeros((max(n, 1), 32), dtype=np.uint8)
        buf = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, dtype=np.uint8)
        _check(self.lib.zkb_sha256_batch(self.handle, buf.ctypes.data, len(data), offs.ctypes.data,
                                         lens.ctypes.data, n, out.ctypes.data), "zkb_sha256_batch")
        return [out[i].tobytes() for i in range(n)]

    def rsa_verify_batch(self, keys: Sequence[bytes], digests: Sequence[bytes], sigs: Sequence[bytes]) -> List[int]:
        n = len(keys)
        kp = (C.c_char_p * max(1, n))(*keys)
        kl = (C.c_size_t * max(1, n))(*[len(k) for k in keys])
        sp = (C.c_char_p * max(1, n))(*sigs)
        sl = (C.c_size_t * max(1, n))(*[len(s) for s in sigs])
        dg = b"".join(digests)
        ok = np.zeros(max(1, n), dtype=np.uint8)
        _check(self.lib.zkb_rsa_verify_batch(self.handle, C.cast(kp, C.c_void_p), C.cast(kl, C.c_void_p),
                                             C.cast(C.c_char_p(dg), C.c_void_p), C.cast(sp, C.c_void_p),
                                             C.cast(sl, C.c_void_p), n, ok.ctypes.data), "zkb_rsa_verify_batch")
        return [int(x) for x in ok[:n]]

    def dfa_scan_batch(self, dfa: DFA, haystacks: Sequence[bytes], qp: bool = False) -> np.ndarray:
        n = len(haystacks)
        lens = np.array([len(h) for h in haystacks], dtype=np.uint32)
        offs = np.zeros(n, dtype=np.uint64)
        pad = [(len(h) + 15) // 16 * 16 for h in haystacks]
        cur = 0
        chunks = []
        for i, h in enumerate(haystacks):
            offs[i] = cur
            chunks.append(h + b"\0" * (pad[i] - len(h)))
            cur += pad[i]
        data = b"".join(chunks) + b"\0" * 16
        f, b = bytes(dfa.fwd), bytes(dfa.bwd)
        view = _DfaView(C.cast(C.c_char_p(f), C.c_void_p).value, len(f), C.cast(C.c_char_p(b), C.c_void_p).value, len(b))
        out = np.zeros((max(1, n), 4), dtype=np.uint32)
        buf = np.frombuffer(data, dtype=np.uint8)
        _check(self.lib.zkb_dfa_scan_batch(self.handle, C.byref(view), buf.ctypes.data, len(data), offs.ctypes.data,
                                           lens.ctypes.data, n, 1 if qp else 0, out.ctypes.data), "zkb_dfa_scan_batch")
        return out[:n]

    def int_pipe_peaks(self) -> dict:
        o = (C.c_double * 8)()
        _check(self.lib.zkb_int_pipe_peaks(self.handle, C.byref(o)), "zkb_int_pipe_peaks")
        return {"imad_wide_gops": o[0], "iadd3_gops": o[1], "lop3_gops": o[2], "shf_gops": o[3],
                "sm_clock_mhz": o[4], "sm_count": int(o[5])}


def _email_output(email: Email, r) -> EmailVerifierOutput:
    ext: List[str] = []
    for x in email.external_inputs:  # core/src/circuits.rs:18-27
        if x.value is None:
            raise VerificationPanic(4)
        ext += [x.name, x.value]
    return EmailVerifierOutput(bytes(r["from_domain_hash"]), bytes(r["public_key_hash"]), ext)


def canonicalize_signed_email(raw_email: bytes, now_unix: int = 0) -> Tuple[bytes, bytes]:
    """cfdkim::canonicalize_signed_email as the reference calls it (core/src/circuits.rs:34-35):
    (header preimage, canonical body) of the first valid DKIM-Signature header.  Host-only."""
    L = load_library()
    h, b = C.c_void_p(), C.c_void_p()
    hl, bl = C.c_size_t(), C.c_size_t()
    detail = C.c_int()
    rc = L.zkb_host_canonicalize(raw_email, len(raw_email), now_unix, C.byref(h), C.byref(hl), C.byref(b),
                                 C.byref(bl), C.byref(detail))
    if rc:
        raise VerificationPanic(1 if detail.value == 1 else 5, detail.value)
    try:
        return C.string_at(h, hl.value), C.string_at(b, bl.value)
    finally:
        L.zkb_free(h)
        L.zkb_free(b)


_default_engine: Optional[Engine] = None


def default_engine() -> Engine:
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine()
    return _default_engine


def verify_email(email: Email) -> EmailVerifierOutput:
    """zkemail_core::verify_email (core/src/circuits.rs:9)."""
    return default_engine().verify_email(email)


def verify_email_with_regex(inp: EmailWithRegex) -> EmailWithRegexVerifierOutput:
    """zkemail_core::verify_email_with_regex (core/src/circuits.rs:31)."""
    return default_engine().verify_email_with_regex(inp)
