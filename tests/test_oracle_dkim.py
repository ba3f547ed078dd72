"""The oracle's DKIM path against mail signed by the independent Python signer (synth.py, written
from RFC 6376) and by the C/OpenSSL generator (workload/zk_gen.c): a signature either verifies or it
does not, so agreement of three independent implementations pins canonicalisation, header
selection and the RSA step."""
import hashlib

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle
import zkemail_rs_b200 as z
from zkemail_rs_b200 import synth
from tests.util import NOW, key_pool, mixed_emails


def test_positive_and_negative_classes():
    emails, labels = mixed_emails(seed=1)
    res = oracle.verify_batch(emails, now=NOW)
    for e, r, lab in zip(emails, res, labels):
        if lab == "pos":
            assert r["status"] == 0 and r["dkim_detail"] == 0 and r["rsa_ok"] == 1 and r["bh_ok"] == 1
            assert r["from_domain_hash"] == hashlib.sha256(e.from_domain.encode()).digest()
            assert r["public_key_hash"] == hashlib.sha256(e.public_key.key).digest()
            hdr, body = oracle.canonicalize_signed_email(e.raw_email, NOW)
            assert r["body_hash"] == hashlib.sha256(body).digest()
            assert r["header_hash"] == hashlib.sha256(hdr).digest()
        else:
            assert r["status"] == 3
    details = {lab: r["dkim_detail"] for lab, r in zip(labels, res)}
    assert details["body_flip"] == 11 and details["bh_flip"] == 11
    assert details["sig_flip"] == 13 and details["wrong_key"] == 13 and details["header_flip"] == 13
    assert details["domain_mismatch"] == 1 and details["missing_tag"] == 3


def test_c_generator_agrees_with_oracle():
    import workload as gen
    kp = gen.KeyPool(3, 2, threads=4)
    mp = gen.MailPool(kp, 240, np.random.default_rng(3).integers(0, 6000, size=240), neg_fraction=0.1,
                      token=True, qp_percent=30, threads=4)
    ems = [mp.email(i) for i in range(mp.n)]
    res = oracle.verify_batch(ems, now=NOW, threads=4)
    for i, r in enumerate(res):
        assert (r["status"] == 0) == bool(mp.expected_ok()[i]), (i, r["status"], r["dkim_detail"], int(mp.neg_kind[i]))
    fast = oracle.verify_batch(ems, now=NOW, threads=4, use_openssl=True)
    assert fast == res  # libcrypto-backed primitives (the timed CPU baseline) give identical records
    kp.close()


def test_multiple_signatures_first_pass_wins():
    rng = np.random.default_rng(5)
    k0, k1 = key_pool()[2048][0], key_pool()[2048][1]
    good = synth.make_email(rng, k0, "a.example.com", idx=1, body_len=300)
    # prepend a signature for another domain (skipped) and a broken one for the same domain
    raw = good.raw_email
    other = synth.make_email(rng, k1, "b.example.com", idx=2, body_len=300).raw_email
    other_sig = other[: other.find(b"Received:")]
    broken = raw[: raw.find(b"Received:")].replace(b"bh=", b"bh=A", 1)
    e = z.Email("a.example.com", other_sig + broken + raw, good.public_key)
    r = oracle.verify_email(e, NOW)
    assert r["status"] == 0 and r["dkim_detail"] == 0
    # only the broken one for this domain -> body hash failure is the last error
    e2 = z.Email("a.example.com", other_sig + broken + raw[raw.find(b"Received:"):], good.public_key)
    r2 = oracle.verify_email(e2, NOW)
    assert r2["status"] == 3 and r2["dkim_detail"] == 11
    # no signature for the domain at all -> neutral
    e3 = z.Email("c.example.com", raw, good.public_key)
    assert oracle.verify_email(e3, NOW)["dkim_detail"] == 1


def test_status_codes_for_panic_sites():
    rng = np.random.default_rng(6)
    k = key_pool()[2048][0]
    e = synth.make_email(rng, k, "a.example.com", idx=1, body_len=100)
    assert oracle.verify_email(z.Email(e.from_domain, b" leading space\r\n\r\n", e.public_key), NOW)["status"] == 1
    assert oracle.verify_email(z.Email(e.from_domain, e.raw_email, z.PublicKey(b"\x30\x00", "rsa")), NOW)["status"] == 2
    assert oracle.verify_email(z.Email(e.from_domain, e.raw_email, z.PublicKey(k.der, "dsa")), NOW)["status"] == 2
    assert oracle.verify_email(z.Email(e.from_domain, e.raw_email, z.PublicKey(b"\x01" * 32, "ed25519")), NOW)["status"] == 9
    assert oracle.verify_email(z.Email(e.from_domain, e.raw_email, z.PublicKey(b"\x01" * 31, "ed25519")), NOW)["status"] == 2
    sha1 = synth.make_email(rng, k, "a.example.com", idx=2, body_len=100, algo="rsa-sha1")
    assert oracle.verify_email(sha1, NOW)["status"] == 9
    expired = synth.make_email(rng, k, "a.example.com", idx=3, body_len=100, extra_tags=" x=1000;")
    assert oracle.verify_email(expired, NOW)["dkim_detail"] == 8
    notyet = synth.make_email(rng, k, "a.example.com", idx=3, body_len=100, extra_tags=f" x={NOW + 10};")
    assert oracle.verify_email(notyet, NOW)["status"] == 0
    badq = synth.make_email(rng, k, "a.example.com", idx=4, body_len=100, extra_tags=" q=dns/other;")
    assert oracle.verify_email(badq, NOW)["dkim_detail"] == 7
    lcut = synth.make_email(rng, k, "a.example.com", idx=5, body_len=100, extra_tags=" l=abc;")
    assert oracle.verify_email(lcut, NOW)["dkim_detail"] == 14


_WS = st.sampled_from([b" ", b"\t", b"  ", b" \t ", b""])
_WORD = st.binary(min_size=0, max_size=12).map(lambda b: bytes(c for c in b if c not in (13, 10)))


@settings(max_examples=300, deadline=None)
@given(st.lists(st.tuples(_WORD, _WS, st.sampled_from([b"\r\n", b"\r\n", b"\n", b"\r", b""])), max_size=12))
def test_relaxed_body_vs_independent_python(parts):
    body = b"".join(w + ws + nl for w, ws, nl in parts)
    got = oracle.canon_body(body, True)
    if b"\n" not in body.replace(b"\r\n", b"") and b"\r" not in body.replace(b"\r\n", b""):
        exp = synth.relaxed_body(body)
        # documented cfdkim quirks: trailing WSP without a final CRLF is kept; "\r\n" stays
        if body.endswith((b" ", b"\t")) or exp == b"\r\n" or got == b"\r\n":
            return
        assert got == exp, (body, got, exp)
    # idempotence (except for the trailing-WSP-without-CRLF quirk, where a CRLF is appended after a kept SP)
    if not body.endswith((b" ", b"\t")):
        assert oracle.canon_body(got, True) == got
