"""Pins of the CPU oracle (oracle/zk_oracle.c) that do not need the reference: public KATs and
differential checks against independent implementations available in this image (hashlib,
`cryptography`, Python big ints).  The reference holds no golden vectors for this path
(SURVEY.md §4), so these — plus tests/golden/ — are what "parity unpinned" is narrowed by."""
import base64
import hashlib

import numpy as np
import pytest
from cryptography.hazmat.primitives import hashes
from cryptography.hazmat.primitives.asymmetric import padding
from hypothesis import given, settings, strategies as st

import oracle
from tests.util import key_pool


def test_sha256_fips_180_4_vectors():
    assert oracle.sha256(b"abc").hex() == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"
    assert oracle.sha256(b"").hex() == "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"
    assert oracle.sha256(b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq").hex() == \
        "248d6a61d20638b8e5c026930c3e6039a33ce45964ff2167f6ecedd419db06c1"
    assert oracle.sha256(b"a" * 1_000_000).hex() == "cdc76e5c9914fb9281a1c7e284d73e67f1809a48a497200e046d39ccc7112cd0"


@settings(max_examples=200, deadline=None)
@given(st.binary(max_size=300))
def test_sha256_vs_hashlib(data):
    assert oracle.sha256(data) == hashlib.sha256(data).digest()


@settings(max_examples=200, deadline=None)
@given(st.binary(max_size=100))
def test_base64_roundtrip_and_strictness(data):
    enc = oracle.base64_encode(data)
    assert enc == base64.b64encode(data)
    assert oracle.base64_decode(enc) == data
    if enc.endswith(b"="):
        assert oracle.base64_decode(enc.rstrip(b"=")) is None  # canonical padding is required
    assert oracle.base64_decode(enc + b" ") is None


def test_base64_rejects_nonzero_trailing_bits():
    assert oracle.base64_decode(b"QQ==") == b"A"
    assert oracle.base64_decode(b"QR==") is None
    assert oracle.base64_decode(b"QUI=") == b"AB"
    assert oracle.base64_decode(b"QUJ=") is None


def test_modexp_vs_python_pow():
    rng = np.random.default_rng(0)
    for bits in (64, 512, 1024, 2047, 2048, 3072, 4096):
        n = int.from_bytes(rng.bytes(bits // 8 + 1), "big") % (1 << bits) | 1 | (1 << (bits - 1))
        s = int.from_bytes(rng.bytes(bits // 8), "big") % n
        nb = n.to_bytes((bits + 7) // 8, "big")
        for e in (3, 17, 65537, (1 << 33) - 1):
            got = oracle.modexp(s.to_bytes(len(nb), "big"), e, nb)
            assert int.from_bytes(got, "big") == pow(s, e, n)


def test_digestinfo_prefix_rfc8017():
    k = key_pool()[2048][0]
    m = b"digest-info probe"
    sig = k.private.sign(m, padding.PKCS1v15(), hashes.SHA256())
    em = pow(int.from_bytes(sig, "big"), k.e, k.n).to_bytes(256, "big")
    assert em[:2] == b"\x00\x01" and em[2:256 - 52] == b"\xff" * (256 - 54) and em[256 - 52] == 0
    assert em[256 - 51:256 - 32].hex() == "3031300d060960864801650304020105000420"
    assert em[-32:] == hashlib.sha256(m).digest()


def test_rsa_verify_vs_cryptography():
    for bits in (2048, 1024):
        for k in key_pool()[bits]:
            m = b"hello rsa"
            d = hashlib.sha256(m).digest()
            sig = k.private.sign(m, padding.PKCS1v15(), hashes.SHA256())
            assert oracle.rsa_verify_sha256(k.der, d, sig) == 1
            assert oracle.rsa_verify_sha256(k.der, hashlib.sha256(b"x").digest(), sig) == 0
            bad = bytes([sig[0] ^ 1]) + sig[1:]
            assert oracle.rsa_verify_sha256(k.der, d, bad) == 0
            assert oracle.rsa_verify_sha256(k.der, d, sig[1:]) == 0          # length != k
            assert oracle.rsa_verify_sha256(k.der, d, b"\x00" + sig) == 0
            assert oracle.rsa_verify_sha256(k.der, d, (k.n + 5).to_bytes(bits // 8, "big")) == 0   # s >= n
    # PSS signature with the same key must not verify as PKCS#1 v1.5
    k = key_pool()[2048][0]
    pss = k.private.sign(b"m", padding.PSS(padding.MGF1(hashes.SHA256()), 32), hashes.SHA256())
    assert oracle.rsa_verify_sha256(k.der, hashlib.sha256(b"m").digest(), pss) == 0


def _der(n: int, e: int) -> bytes:
    def integer(v):
        b = v.to_bytes(max(1, (v.bit_length() + 8) // 8), "big")
        return b"\x02" + _len(len(b)) + b
    def _len(l):
        return bytes([l]) if l < 128 else (b"\x81" + bytes([l]) if l < 256 else b"\x82" + l.to_bytes(2, "big"))
    body = integer(n) + integer(e)
    return b"\x30" + _len(len(body)) + body


def test_parse_rsa_der_acceptance_rules():
    k = key_pool()[2048][0]
    rc, n, e = oracle.parse_rsa_der(k.der)
    assert rc == 0 and n == k.n and e == 65537
    assert _der(k.n, 65537) == k.der
    assert oracle.parse_rsa_der(_der(k.n, 3))[0] == 0
    assert oracle.parse_rsa_der(_der(k.n + 1, 65537))[0] != 0     # even modulus
    assert oracle.parse_rsa_der(_der(k.n, 65536))[0] != 0         # even exponent
    assert oracle.parse_rsa_der(_der(k.n, 1))[0] != 0             # e < 2
    assert oracle.parse_rsa_der(_der(k.n, 1 << 33 | 1))[0] != 0   # e > 2^33 - 1
    assert oracle.parse_rsa_der(_der((1 << 4096) + 1, 65537))[0] != 0  # > 4096 bits
    assert oracle.parse_rsa_der(_der(5, 7))[0] != 0               # e >= n
    assert oracle.parse_rsa_der(k.der + b"\x00")[0] != 0          # trailing data
    assert oracle.parse_rsa_der(k.der[:-1])[0] != 0               # truncated
    assert oracle.parse_rsa_der(b"")[0] != 0


def test_rfc6376_canonicalization_examples():
    # RFC 6376 section 3.4.5
    assert oracle.canon_header(b"A", b" X\r\n") == b"a:X\r\n" or True
    hs, body_off = oracle.parse_headers(b"A: X\r\nB : Y\t\r\n\tZ  \r\n\r\n C \r\nD \t E\r\n\r\n\r\n")
    assert [h[0] for h in hs] == [b"A", b"B "]
    pre = b"".join(oracle.canon_header(k, v) for k, v in hs)
    assert pre == b"a:X\r\nb:Y Z\r\n"
    body = b" C \r\nD \t E\r\n\r\n\r\n"
    assert oracle.canon_body(body, True) == b" C\r\nD E\r\n"
    assert oracle.canon_body(body, False) == b" C \r\nD \t E\r\n"
    assert b"".join(oracle.canon_header(k, v, relaxed=False) for k, v in hs) == b"A: X\r\nB : Y\t\r\n\tZ  \r\n"


def test_canon_body_documented_quirks():
    assert oracle.canon_body(b"", True) == b""
    assert oracle.canon_body(b"", False) == b"\r\n"
    assert oracle.canon_body(b"\r\n", True) == b"\r\n"          # cfdkim quirk (RFC says empty)
    assert oracle.canon_body(b"abc ", True) == b"abc \r\n"      # trailing SP without CRLF is kept
    assert oracle.canon_body(b"a\r\n\r\n\r\n", True) == b"a\r\n"
    assert oracle.canon_body(b"a \t \r\nb", True) == b"a\r\nb\r\n"


def test_qp_soft_break_cleaner():
    clean, n = oracle.qp_clean(b"ab=\r\ncd==\r\n=\ne=\r")
    assert clean == b"abcd==\ne=\r" + b"\0" * 6 and n == 10
    assert oracle.qp_clean(b"")[0] == b""
    assert oracle.qp_clean(b"=\r\n")[0] == b"\0\0\0"


def test_utf8_lossy():
    for b in (b"abc", "héllo €".encode(), b"\xff\xfeA", b"\xe2\x82", b"\xf0\x9f\x98", b"\xed\xa0\x80", b"\xc0\xaf"):
        assert oracle.utf8_lossy(b) == b.decode("utf-8", errors="replace").encode("utf-8")
