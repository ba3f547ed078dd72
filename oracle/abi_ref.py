"""TEST ORACLE (not product code): generic contract-ABI encoder, written from the ABI specification's
head/tail rules, used to check the library's packer for core/src/io.rs:5-53 (SolEmailOutput /
SolEmailWithRegexOutput via alloy-sol-types' SolValue::abi_encode — alloy-sol-types 0.8.x is a Cargo.lock
dependency absent from /root/reference).  Pinned by the three worked examples of the ABI specification
(tests/test_abi_io.py); parity with the Rust crate itself is unpinned (no Rust toolchain here).

Values are tagged tuples: ("uint", int) ("bool", b) ("bytes32", b) ("bytesN", b) ("bytes", b) ("string", str)
("array", [values]) ("tuple", [values])."""
from __future__ import annotations


def _word(n: int) -> bytes:
    return int(n).to_bytes(32, "big")


def _pad_right(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 32)


def is_dynamic(v) -> bool:
    t, x = v
    if t in ("bytes", "string", "array"):
        return True
    if t == "tuple":
        return any(is_dynamic(e) for e in x)
    return False


def enc(v) -> bytes:
    t, x = v
    if t == "uint":
        return _word(x)
    if t == "bool":
        return _word(1 if x else 0)
    if t in ("bytes32", "bytesN"):
        assert len(x) <= 32
        return _pad_right(bytes(x))
    if t == "bytes":
        return _word(len(x)) + _pad_right(bytes(x))
    if t == "string":
        b = x.encode("utf-8") if isinstance(x, str) else bytes(x)
        return _word(len(b)) + _pad_right(b)
    if t == "array":
        return _word(len(x)) + enc_sequence(x)
    if t == "tuple":
        return enc_sequence(x)
    raise ValueError(t)


def enc_sequence(vals) -> bytes:
    """head/tail encoding of a tuple's members (function arguments, struct fields, array elements)."""
    head_size = sum(32 if is_dynamic(v) else len(enc(v)) for v in vals)
    head, tail = b"", b""
    for v in vals:
        if is_dynamic(v):
            head += _word(head_size + len(tail))
            tail += enc(v)
        else:
            head += enc(v)
    return head + tail


def sol_email_output(fdh: bytes, pkh: bytes, external_inputs) -> tuple:
    return ("tuple", [("bytes32", fdh), ("bytes32", pkh), ("array", [("string", s) for s in external_inputs])])


def verification_output_abi_encode(fdh: bytes, pkh: bytes, external_inputs, matches=None) -> bytes:
    """core/src/io.rs:35-45: abi_encode of one value == encoding of the 1-element sequence holding it."""
    email = sol_email_output(fdh, pkh, external_inputs)
    if matches is None:
        return enc_sequence([email])
    return enc_sequence([("tuple", [email, ("array", [("string", s) for s in matches])])])
