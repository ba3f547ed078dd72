"""Solidity-ABI packing of the verifier outputs — mirror of the reference's
core/src/io.rs:18-53 (VerificationOutput, from_parts, abi_encode) and helpers/src/io.rs:6-31
(AbiDecodable::abi_decode), over the library's batch packer (zkb_abi_encode_batch / zkb_abi_decode,
csrc/abi_pack.hpp).  Host code; needs no device."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

from .engine import _check, load_library
from .structs import EmailVerifierOutput, EmailWithRegexVerifierOutput


class AbiDecodeError(ValueError):
    """helpers/src/io.rs:7-9: abi_decode returned Err (not a canonical encoding of either struct)."""


class _Str(C.Structure):
    _fields_ = [("s", C.c_char_p), ("len", C.c_size_t)]


class _Span(C.Structure):
    _fields_ = [("off", C.c_uint64), ("len", C.c_uint64)]


class _OutputView(C.Structure):
    _fields_ = [("from_domain_hash", C.c_char_p), ("public_key_hash", C.c_char_p),
                ("external_inputs", C.POINTER(_Str)), ("n_external_inputs", C.c_size_t),
                ("matches", C.POINTER(_Str)), ("n_matches", C.c_size_t), ("with_regex", C.c_int)]


class _Decoded(C.Structure):
    _fields_ = [("with_regex", C.c_int32), ("from_domain_hash", C.c_uint8 * 32), ("public_key_hash", C.c_uint8 * 32),
                ("n_external_inputs", C.c_uint32), ("n_matches", C.c_uint32)]


def _lib():
    L = load_library()
    if not getattr(L, "_zkb_abi_bound", False):
        L.zkb_abi_encode_batch.argtypes = [C.POINTER(_OutputView), C.c_size_t, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.zkb_abi_decode.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(_Decoded), C.POINTER(C.c_void_p)]
        L._zkb_abi_bound = True
    return L


@dataclass
class VerificationOutput:
    """core/src/io.rs:18-25: EmailOnly(email) when matches is None, WithRegex{email, matches} otherwise."""
    email: EmailVerifierOutput
    matches: Optional[List[str]] = None

    @classmethod
    def from_parts(cls, email: EmailVerifierOutput, matches: Optional[List[str]]) -> "VerificationOutput":
        return cls(email, matches)  # core/src/io.rs:28-33

    @classmethod
    def from_output(cls, out) -> "VerificationOutput":
        if isinstance(out, EmailWithRegexVerifierOutput):
            return cls(out.email, list(out.regex_matches))
        return cls(out, None)

    def abi_encode(self) -> bytes:  # core/src/io.rs:35-45
        return abi_encode_batch([self])[0]

    @classmethod
    def abi_decode(cls, data: bytes) -> "VerificationOutput":  # helpers/src/io.rs:12-31
        return abi_decode(data)


def abi_encode_batch(outputs: Sequence[VerificationOutput], threads: int = 0) -> List[bytes]:
    """VerificationOutput::abi_encode for a whole batch in one library call."""
    n = len(outputs)
    if n == 0:
        return []
    views = (_OutputView * n)()
    keep = []
    for i, o in enumerate(outputs):
        fdh, pkh = bytes(o.email.from_domain_hash), bytes(o.email.public_key_hash)
        if len(fdh) != 32 or len(pkh) != 32:  # core/src/io.rs:49-50 try_into().unwrap()
            raise ValueError("hashes must be 32 bytes")
        ext = [s.encode("utf-8") for s in o.email.external_inputs]
        mt = [s.encode("utf-8") for s in (o.matches or [])]
        ea = (_Str * max(1, len(ext)))(*[_Str(b, len(b)) for b in ext])
        ma = (_Str * max(1, len(mt)))(*[_Str(b, len(b)) for b in mt])
        keep.append((fdh, pkh, ext, mt, ea, ma))
        v = views[i]
        v.from_domain_hash, v.public_key_hash = fdh, pkh
        v.external_inputs, v.n_external_inputs = ea, len(ext)
        v.matches, v.n_matches = ma, len(mt)
        v.with_regex = 0 if o.matches is None else 1
    blob = C.c_void_p()
    offs = (C.c_uint64 * (n + 1))()
    L = _lib()
    _check(L.zkb_abi_encode_batch(views, n, threads, C.byref(blob), offs), "zkb_abi_encode_batch")
    try:
        raw = C.string_at(blob.value, offs[n])
    finally:
        L.zkb_free(blob)
    return [raw[offs[i]:offs[i + 1]] for i in range(n)]


def abi_decode(data: bytes) -> VerificationOutput:
    L = _lib()
    dec = _Decoded()
    spans = C.c_void_p()
    rc = L.zkb_abi_decode(data, len(data), C.byref(dec), C.byref(spans))
    if rc != 0:
        raise AbiDecodeError("not a canonical SolEmailOutput / SolEmailWithRegexOutput encoding")
    try:
        n = dec.n_external_inputs + dec.n_matches
        sp = C.cast(spans, C.POINTER(_Span))
        strs = [data[sp[i].off:sp[i].off + sp[i].len].decode("utf-8") for i in range(n)]
    finally:
        L.zkb_free(spans)
    email = EmailVerifierOutput(bytes(dec.from_domain_hash), bytes(dec.public_key_hash), strs[:dec.n_external_inputs])
    return VerificationOutput(email, strs[dec.n_external_inputs:] if dec.with_regex else None)
