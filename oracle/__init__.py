"""ctypes binding of the CPU ORACLE (oracle/zk_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package.  "parity unpinned" w.r.t. the
Rust binary (see zk_oracle.h)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libzk_oracle.so")
ZO_MAX_PARTS = 16

STATUS_NAMES = {
    0: "OK", 1: "MAIL_PARSE", 2: "KEY", 3: "DKIM_FAIL", 4: "NULL_EXTERNAL", 5: "CANONICALIZE",
    6: "REGEX_HEADER", 7: "REGEX_BODY", 8: "BAD_DFA", 9: "UNSUPPORTED",
}


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("zk_oracle.c", "zk_oracle.h")]
    src = [s for s in src if os.path.exists(s)]
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class _Part(C.Structure):
    _fields_ = [("match_count", C.c_uint32), ("start", C.c_uint32), ("end", C.c_uint32),
                ("captures_ok", C.c_uint32)]


class Result(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("dkim_detail", C.c_int32),
        ("body_hash", C.c_uint8 * 32), ("header_hash", C.c_uint8 * 32),
        ("from_domain_hash", C.c_uint8 * 32), ("public_key_hash", C.c_uint8 * 32),
        ("bh_ok", C.c_uint8), ("rsa_ok", C.c_uint8), ("pad", C.c_uint8 * 2),
        ("n_parts", C.c_uint32), ("parts", _Part * ZO_MAX_PARTS),
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


class _RegexPart(C.Structure):
    _fields_ = [("fwd", C.c_void_p), ("fwd_len", C.c_size_t), ("bwd", C.c_void_p),
                ("bwd_len", C.c_size_t), ("captures", C.POINTER(C.c_char_p)),
                ("n_captures", C.c_size_t)]


class _Email(C.Structure):
    _fields_ = [("from_domain", C.c_char_p), ("from_domain_len", C.c_size_t),
                ("raw_email", C.c_void_p), ("raw_len", C.c_size_t),
                ("key", C.c_void_p), ("key_len", C.c_size_t), ("key_type", C.c_char_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.zo_sha256.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        L.zo_base64_encode.restype = C.c_size_t
        L.zo_base64_decode.restype = C.c_long
        L.zo_modexp.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64, C.c_char_p, C.c_size_t, C.c_char_p]
        L.zo_parse_rsa_der.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.POINTER(C.c_size_t),
                                       C.POINTER(C.c_uint64)]
        L.zo_rsa_verify_sha256.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_size_t]
        for f in ("zo_canon_body_relaxed", "zo_canon_body_simple"):
            getattr(L, f).argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
            getattr(L, f).restype = C.c_size_t
        for f in ("zo_canon_header_relaxed", "zo_canon_header_simple"):
            getattr(L, f).argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p]
            getattr(L, f).restype = C.c_size_t
        L.zo_parse_headers.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32), C.c_size_t,
                                       C.POINTER(C.c_size_t)]
        L.zo_parse_headers.restype = C.c_long
        L.zo_utf8_lossy.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        L.zo_utf8_lossy.restype = C.c_size_t
        L.zo_qp_clean.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        L.zo_qp_clean.restype = C.c_size_t
        L.zo_canonicalize_signed_email.argtypes = [
            C.c_char_p, C.c_size_t, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
            C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.zo_free.argtypes = [C.c_void_p]
        L.zo_dfa_find_iter.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p,
                                       C.c_size_t, C.POINTER(C.c_uint32), C.c_size_t]
        L.zo_dfa_find_iter.restype = C.c_long
        L.zo_verify_email.argtypes = [C.POINTER(_Email), C.c_int64, C.POINTER(Result)]
        L.zo_verify_email_with_regex.argtypes = [
            C.POINTER(_Email), C.POINTER(_RegexPart), C.c_size_t, C.POINTER(_RegexPart), C.c_size_t,
            C.c_int64, C.POINTER(Result)]
        L.zo_verify_batch_mt.argtypes = [
            C.POINTER(_Email), C.c_size_t, C.POINTER(_RegexPart), C.c_size_t, C.POINTER(_RegexPart),
            C.c_size_t, C.c_int64, C.c_int, C.c_int, C.POINTER(Result)]
        _lib = L
    return _lib


# ------------------------------------------------------------------ thin helpers


def sha256(data: bytes) -> bytes:
    out = C.create_string_buffer(32)
    lib().zo_sha256(data, len(data), out)
    return out.raw


def base64_encode(data: bytes) -> bytes:
    out = C.create_string_buffer(4 * ((len(data) + 2) // 3) + 4)
    n = lib().zo_base64_encode(data, len(data), out)
    return out.raw[:n]


def base64_decode(data: bytes) -> Optional[bytes]:
    out = C.create_string_buffer(len(data) + 4)
    n = lib().zo_base64_decode(data, len(data), out)
    return None if n < 0 else out.raw[:n]


def modexp(base: bytes, e: int, mod: bytes) -> bytes:
    out = C.create_string_buffer(len(mod))
    rc = lib().zo_modexp(base, len(base), e, mod, len(mod), out)
    assert rc == 0
    return out.raw


def parse_rsa_der(der: bytes) -> Tuple[int, Optional[int], Optional[int]]:
    nb = C.create_string_buffer(1024)
    nl = C.c_size_t()
    e = C.c_uint64()
    rc = lib().zo_parse_rsa_der(der, len(der), nb, C.byref(nl), C.byref(e))
    if rc:
        return rc, None, None
    return 0, int.from_bytes(nb.raw[: nl.value], "big"), e.value


def rsa_verify_sha256(der: bytes, digest: bytes, sig: bytes) -> int:
    return lib().zo_rsa_verify_sha256(der, len(der), digest, sig, len(sig))


def canon_body(body: bytes, relaxed: bool = True) -> bytes:
    out = C.create_string_buffer(len(body) + 8)
    f = lib().zo_canon_body_relaxed if relaxed else lib().zo_canon_body_simple
    n = f(body, len(body), out)
    return out.raw[:n]


def canon_header(key: bytes, value: bytes, relaxed: bool = True) -> bytes:
    out = C.create_string_buffer(2 * len(key) + len(value) + 16)
    f = lib().zo_canon_header_relaxed if relaxed else lib().zo_canon_header_simple
    n = f(key, len(key), value, len(value), out)
    return out.raw[:n]


def parse_headers(raw: bytes):
    cap = raw.count(b"\n") + 2
    quads = (C.c_uint32 * (4 * cap))()
    body_off = C.c_size_t()
    n = lib().zo_parse_headers(raw, len(raw), quads, cap, C.byref(body_off))
    if n < 0:
        return None
    hs = []
    for i in range(n):
        ko, kl, vo, vl = quads[4 * i : 4 * i + 4]
        hs.append((raw[ko : ko + kl], raw[vo : vo + vl]))
    return hs, body_off.value


def utf8_lossy(b: bytes) -> bytes:
    out = C.create_string_buffer(3 * len(b) + 1)
    n = lib().zo_utf8_lossy(b, len(b), out)
    return out.raw[:n]


def qp_clean(b: bytes) -> Tuple[bytes, int]:
    out = C.create_string_buffer(len(b) + 1)
    n = lib().zo_qp_clean(b, len(b), out)
    return out.raw[: len(b)], n


def canonicalize_signed_email(raw: bytes, now: int = 1):
    hp, bp = C.c_void_p(), C.c_void_p()
    hl, bl = C.c_size_t(), C.c_size_t()
    rc = lib().zo_canonicalize_signed_email(raw, len(raw), now, C.byref(hp), C.byref(hl),
                                            C.byref(bp), C.byref(bl))
    if rc:
        return None
    h = C.string_at(hp, hl.value)
    b = C.string_at(bp, bl.value)
    lib().zo_free(hp)
    lib().zo_free(bp)
    return h, b


def dfa_find_iter(fwd: bytes, bwd: bytes, hay: bytes, cap: int = 64):
    spans = (C.c_uint32 * (2 * cap))()
    n = lib().zo_dfa_find_iter(fwd, len(fwd), bwd, len(bwd), hay, len(hay), spans, cap)
    if n < 0:
        return None
    return n, [(spans[2 * i], spans[2 * i + 1]) for i in range(min(n, cap))]


class _Keep:
    """Keeps ctypes buffers alive for the lifetime of a marshalled batch."""

    def __init__(self):
        self.refs = []

    def buf(self, b: bytes):
        a = C.create_string_buffer(bytes(b), len(b)) if len(b) else C.create_string_buffer(1)
        self.refs.append(a)
        return C.cast(a, C.c_void_p)


def _marshal_emails(emails, keep: _Keep):
    arr = (_Email * max(1, len(emails)))()
    for i, e in enumerate(emails):
        dom = e.from_domain.encode()
        keep.refs.append(dom)
        arr[i].from_domain = dom
        arr[i].from_domain_len = len(dom)
        arr[i].raw_email = keep.buf(e.raw_email)
        arr[i].raw_len = len(e.raw_email)
        arr[i].key = keep.buf(e.public_key.key)
        arr[i].key_len = len(e.public_key.key)
        kt = e.public_key.key_type.encode()
        keep.refs.append(kt)
        arr[i].key_type = kt
    return arr


def _marshal_parts(parts, keep: _Keep):
    """parts: list of CompiledRegex-like (verify_re.fwd/bwd, captures)"""
    if parts is None:
        return None, 0
    arr = (_RegexPart * max(1, len(parts)))()
    for i, p in enumerate(parts):
        arr[i].fwd = keep.buf(p.verify_re.fwd)
        arr[i].fwd_len = len(p.verify_re.fwd)
        arr[i].bwd = keep.buf(p.verify_re.bwd)
        arr[i].bwd_len = len(p.verify_re.bwd)
        if p.captures is None:
            arr[i].captures = None
            arr[i].n_captures = 0
        else:
            cs = (C.c_char_p * max(1, len(p.captures)))()
            for j, s in enumerate(p.captures):
                cs[j] = s.encode()
            keep.refs.append(cs)
            arr[i].captures = cs
            arr[i].n_captures = len(p.captures)
    return arr, len(parts)


def verify_email(email, now: int = 1) -> dict:
    keep = _Keep()
    arr = _marshal_emails([email], keep)
    r = Result()
    lib().zo_verify_email(arr, now, C.byref(r))
    return r.as_dict()


def verify_email_with_regex(ewr, now: int = 1) -> dict:
    keep = _Keep()
    arr = _marshal_emails([ewr.email], keep)
    hp, nh = _marshal_parts(ewr.regex_info.header_parts, keep)
    bp, nb = _marshal_parts(ewr.regex_info.body_parts, keep)
    r = Result()
    lib().zo_verify_email_with_regex(arr, hp, nh, bp, nb, now, C.byref(r))
    return r.as_dict()


def verify_batch(emails, header_parts=None, body_parts=None, now: int = 1, threads: int = 1,
                 use_openssl: bool = False) -> List[dict]:
    keep = _Keep()
    arr = _marshal_emails(emails, keep)
    hp, nh = _marshal_parts(header_parts, keep)
    bp, nb = _marshal_parts(body_parts, keep)
    out = (Result * max(1, len(emails)))()
    rc = lib().zo_verify_batch_mt(arr, len(emails), hp, nh, bp, nb, now, threads,
                                  1 if use_openssl else 0, out)
    assert rc == 0
    return [out[i].as_dict() for i in range(len(emails))]
