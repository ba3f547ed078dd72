"""GPU parity on the edge cases the domain has: empty and ragged batches,
extreme sizes, every status code, several signatures per message, odd keys.  Everything goes
through the C ABI; the oracle is the checker."""
import hashlib

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle
import zkemail_rs_b200 as z
from zkemail_rs_b200 import synth
from zkemail_rs_b200.structs import RegexInfo, RegexPattern, CompiledRegex
from tests.util import NOW, assert_records_equal, key_pool, mixed_emails

pytestmark = pytest.mark.gpu


def _check(engine, emails, labels=None):
    got = engine.verify_batch(emails)
    exp = oracle.verify_batch(emails, now=NOW)
    for i, (g, e) in enumerate(zip(got, exp)):
        assert_records_equal(g, e, labels[i] if labels else i)
    return got


def test_empty_and_single_batches(engine):
    assert len(engine.verify_batch([])) == 0
    rng = np.random.default_rng(1)
    e = synth.make_email(rng, key_pool()[2048][0], "a.example.com", idx=0, body_len=10)
    _check(engine, [e])
    assert engine.sha256_batch([]) == []
    assert engine.sha256_batch([b""]) == [hashlib.sha256(b"").digest()]


def test_ragged_sizes_and_large_body(engine):
    rng = np.random.default_rng(2)
    keys = key_pool()
    sizes = [0, 1, 2, 3, 61, 62, 63, 64, 65, 117, 118, 119, 120, 121, 1023, 1024, 1025, 65536, 300_000, 1_200_000]
    emails = [synth.make_email(rng, keys[2048 if i % 3 else 1024][i % 2], f"r{i % 4}.example.com", idx=i, body_len=s)
              for i, s in enumerate(sizes)]
    got = _check(engine, emails)
    assert all(int(g["status"]) == 0 for g in got)


def test_every_status_code(engine):
    rng = np.random.default_rng(3)
    k = key_pool()[2048][0]
    base = synth.make_email(rng, k, "s.example.com", idx=1, body_len=200)
    P = z.PublicKey
    emails = [
        base,                                                                        # 0 OK
        z.Email(base.from_domain, b" leading space\r\n\r\nbody", base.public_key),   # 1 MAIL_PARSE
        z.Email(base.from_domain, b"A: b\r\n\rX", base.public_key),                  # 1 (lone CR after headers)
        z.Email(base.from_domain, base.raw_email, P(b"\x30\x00", "rsa")),            # 2 KEY
        z.Email(base.from_domain, base.raw_email, P(k.der, "dsa")),                  # 2
        z.Email(base.from_domain, base.raw_email, P(k.der[:-1], "rsa")),             # 2 (truncated DER)
        z.Email(base.from_domain, base.raw_email, P(b"\x01" * 32, "ed25519")),       # 9 UNSUPPORTED
        z.Email(base.from_domain, base.raw_email, P(b"\x01" * 31, "ed25519")),       # 2
        synth.make_email(rng, k, "s.example.com", idx=2, body_len=50, algo="rsa-sha1"),        # 9
        synth.make_email(rng, k, "s.example.com", idx=3, body_len=50, algo="ed25519-sha256"),  # 3 / ALGO_KEY_MISMATCH
        synth.make_email(rng, k, "s.example.com", idx=4, body_len=50, algo="rsa-md5"),         # 3 / HASH_ALGO
        synth.make_email(rng, k, "s.example.com", idx=5, body_len=50, canon="relaxed/strict"), # 3 / CANON_TYPE
        synth.make_email(rng, k, "s.example.com", idx=6, body_len=50, extra_tags=" x=1000;"),  # 3 / EXPIRED
        synth.make_email(rng, k, "s.example.com", idx=7, body_len=50, extra_tags=" q=dns/other;"),
        synth.make_email(rng, k, "s.example.com", idx=8, body_len=50, extra_tags=" l=abc;"),
        synth.make_email(rng, k, "s.example.com", idx=9, body_len=50, extra_tags=" i=@other.org;"),
        synth.make_email(rng, k, "s.example.com", idx=10, body_len=50, h=("to", "subject")),   # From not signed
        z.Email(base.from_domain, b"no headers at all", base.public_key),
        z.Email(base.from_domain, b"", base.public_key),
        z.Email("", base.raw_email, base.public_key),
        z.Email(base.from_domain.upper(), base.raw_email, base.public_key),          # domain compare is case-insensitive
    ]
    got = _check(engine, emails)
    st_ = [int(g["status"]) for g in got]
    assert st_[:9] == [0, 1, 1, 2, 2, 2, 9, 2, 9]
    assert int(got[9]["dkim_detail"]) == 15 and int(got[10]["dkim_detail"]) == 10 and int(got[11]["dkim_detail"]) == 9
    assert st_[-1] == 0


def test_l_tag_and_truncated_body(engine):
    rng = np.random.default_rng(4)
    k = key_pool()[2048][1]
    # l=100 covers exactly the signed (canonical-stable, CRLF-terminated) body: text appended after
    # signing must not break the signature; l=99 truncates the hashed body and must
    headers = synth.default_headers(rng, "l.example.com", 1)
    body = synth.synth_body(rng, 100)
    raw = synth.sign_email(headers, body, k, "l.example.com", extra_tags=" l=100;")
    e1 = z.Email("l.example.com", raw + b"appended after signing\r\n", z.PublicKey(k.der, "rsa"))
    e2 = z.Email("l.example.com", raw.replace(b" l=100;", b" l=99;"), z.PublicKey(k.der, "rsa"))
    got = _check(engine, [e1, e2])
    assert int(got[0]["status"]) == 0 and int(got[1]["status"]) == 3
    # regex parts see the TRUNCATED canonical body (the appended text is not part of the haystack), captures included
    from tests.util import contiguous_views
    body3 = b"Transaction ID: ZX81\r\n" + b"a" * 36 + b"\r\n" + b"b" * 38 + b"\r\n"
    raw3 = synth.sign_email(headers, body3, k, "l.example.com", extra_tags=" l=100;")
    e3 = z.Email("l.example.com", raw3 + b"Transaction ID: LATE99 appended after signing\r\n", z.PublicKey(k.der, "rsa"))
    for pat, caps, want in ((r"Transaction ID: [A-Z0-9]+", ["ZX81"], 0), (r"Transaction ID: [A-Z0-9]+", ["LATE99"], 7), (r"appended", None, 7)):
        info = RegexInfo(None, [CompiledRegex(z.compile_regex(pat), caps)])
        ex = oracle.verify_batch([e3], None, info.body_parts, now=NOW)
        g = engine.verify_with_regex_batch([e3], info)
        assert_records_equal(g[0], ex[0], (pat, caps))
        assert int(g[0]["status"]) == want, (pat, caps)
        buf, views = contiguous_views([e3])
        engine.register_host(buf)
        try:
            rs = z.RegexSet(engine, info)
            assert_records_equal(engine.verify_views(views, rs)[0], ex[0], ("registered", pat, caps))
        finally:
            engine.unregister_host(buf)


def test_multiple_signatures(engine):
    rng = np.random.default_rng(5)
    k0, k1 = key_pool()[2048][0], key_pool()[2048][1]
    good = synth.make_email(rng, k0, "a.example.com", idx=1, body_len=300)
    raw = good.raw_email
    other = synth.make_email(rng, k1, "b.example.com", idx=2, body_len=300).raw_email
    other_sig = other[: other.find(b"Received:")]
    broken = raw[: raw.find(b"Received:")].replace(b"bh=", b"bh=A", 1)
    simple = synth.make_email(rng, k0, "a.example.com", idx=3, body_len=300, canon="simple/simple")
    emails = [
        z.Email("a.example.com", other_sig + broken + raw, good.public_key),      # third signature passes
        z.Email("a.example.com", other_sig + broken + raw[raw.find(b"Received:"):], good.public_key),
        z.Email("c.example.com", raw, good.public_key),                           # neutral
        z.Email("a.example.com", broken + broken + broken + raw, good.public_key),
        simple,
    ]
    got = _check(engine, emails)
    assert [int(g["status"]) for g in got] == [0, 3, 3, 0, 0]
    # with regex: the haystacks come from the FIRST valid signature (other domain), not the verifying one
    info = RegexInfo([CompiledRegex(z.compile_regex(r"d=b\.example\.com"), None)], None)
    g2 = engine.verify_with_regex_batch(emails[:1], info)
    e2 = oracle.verify_batch(emails[:1], info.header_parts, None, now=NOW)
    assert_records_equal(g2[0], e2[0], "haystack-from-first-valid-signature")
    assert int(g2[0]["status"]) == 0


def _forge(nbits, e, rng):
    from tests.test_emu_kernels import _prime, _der
    import math
    while True:
        p, q = _prime(nbits // 2, rng), _prime(nbits - nbits // 2, rng)
        n, phi = p * q, (p - 1) * (q - 1)
        if n.bit_length() == nbits and math.gcd(e, phi) == 1:
            return n, pow(e, -1, phi), _der(n, e)


@pytest.mark.parametrize("nbits,e", [(2047, 65537), (1536, 65537), (2048, 3), (1000, 17), (3072, 65537), (4096, 65537), (768, 65537)])
def test_odd_keys_through_the_rsa_entry_point(engine, nbits, e):
    rng = np.random.default_rng(nbits + e)
    n, d, der = _forge(nbits, e, rng)
    k = (nbits + 7) // 8
    ks, ds, ss = [], [], []
    for i in range(8):
        h = hashlib.sha256(b"m%d" % i).digest()
        em = b"\x00\x01" + b"\xff" * (k - 54) + b"\x00" + bytes.fromhex("3031300d060960864801650304020105000420") + h
        s = pow(int.from_bytes(em, "big"), d, n)
        if i % 4 == 1:
            s ^= 1 << 9
        if i % 4 == 2:
            h = hashlib.sha256(b"other").digest()
        ks.append(der); ds.append(h); ss.append(s.to_bytes(k, "big"))
    got = engine.rsa_verify_batch(ks, ds, ss)
    exp = [oracle.rsa_verify_sha256(a, b, c) for a, b, c in zip(ks, ds, ss)]
    assert got == exp and sum(got) == 4


_NAME = st.sampled_from([b"From", b"from", b"To", b"Subject", b"Date", b"X-Test", b"Cc"])
_VAL = st.lists(st.sampled_from([b"a", b"B", b" ", b"\t", b"  ", b"\r\n ", b"x@y.z", b";", b"=", b"\xc3\xa9", b"\xff"]), max_size=8).map(b"".join)
_BODY = st.lists(st.sampled_from([b"line", b" ", b"\t", b"\r\n", b"\n", b"\r", b"=\r\n", b"  ", b"text text", b""]), max_size=14).map(b"".join)


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(st.lists(st.tuples(st.lists(st.tuples(_NAME, _VAL), min_size=1, max_size=6), _BODY,
                          st.sampled_from(["relaxed/relaxed", "simple/simple", "relaxed/simple", "simple/relaxed"])),
                min_size=1, max_size=8))
def test_dirty_mail_batches_vs_oracle(engine, mails):
    """Randomly dirty (but signed) mail: folds, tabs, bare CR/LF, non-ASCII, odd canonicalisation."""
    rng = np.random.default_rng(7)
    k = key_pool()[2048][0]
    emails = []
    for headers, body, canon in mails:
        hs = [(n.decode("latin1"), v.rstrip(b"\r\n\t ").decode("latin1")) for n, v in headers]
        if not any(n.lower() == "from" for n, _ in hs):
            hs.append(("From", "x@d.example.com"))
        try:
            raw = synth.sign_email(hs, body, k, "d.example.com", canon=canon)
        except Exception:
            continue
        emails.append(z.Email("d.example.com", raw, z.PublicKey(k.der, "rsa")))
    if emails:
        _check(engine, emails)


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(st.lists(st.tuples(st.lists(st.tuples(_NAME, _VAL), min_size=1, max_size=6), _BODY,
                          st.sampled_from(["relaxed/relaxed", "simple/simple", "relaxed/simple", "simple/relaxed"])),
                min_size=1, max_size=8))
def test_dirty_mail_batches_registered_memory_vs_oracle(engine, mails):
    """The same dirty mail through the device front end (registered memory)."""
    from tests.util import contiguous_views
    k = key_pool()[2048][0]
    emails = []
    for headers, body, canon in mails:
        hs = [(n.decode("latin1"), v.rstrip(b"\r\n\t ").decode("latin1")) for n, v in headers]
        if not any(n.lower() == "from" for n, _ in hs):
            hs.append(("From", "x@d.example.com"))
        try:
            raw = synth.sign_email(hs, body, k, "d.example.com", canon=canon)
        except Exception:
            continue
        emails.append(z.Email("d.example.com", raw, z.PublicKey(k.der, "rsa")))
    if not emails:
        return
    buf, views = contiguous_views(emails)
    engine.register_host(buf)
    try:
        got = engine.verify_views(views)
    finally:
        engine.unregister_host(buf)
    exp = oracle.verify_batch(emails, now=NOW)
    for i, (g, e) in enumerate(zip(got, exp)):
        assert_records_equal(g, e, i)


def test_status_codes_registered_memory(engine):
    """Every status code again, with the messages in registered memory (device front end + fallback)."""
    from tests.util import contiguous_views
    rng = np.random.default_rng(3)
    k = key_pool()[2048][0]
    base = synth.make_email(rng, k, "s.example.com", idx=1, body_len=200)
    P = z.PublicKey
    emails = [
        base,
        z.Email(base.from_domain, b" leading space\r\n\r\nbody", base.public_key),
        z.Email(base.from_domain, b"A: b\r\n\rX", base.public_key),
        z.Email(base.from_domain, base.raw_email, P(b"\x30\x00", "rsa")),
        z.Email(base.from_domain, base.raw_email, P(b"\x01" * 32, "ed25519")),
        synth.make_email(rng, k, "s.example.com", idx=2, body_len=50, algo="rsa-sha1"),
        synth.make_email(rng, k, "s.example.com", idx=3, body_len=50, algo="ed25519-sha256"),
        synth.make_email(rng, k, "s.example.com", idx=5, body_len=50, canon="relaxed/strict"),
        synth.make_email(rng, k, "s.example.com", idx=6, body_len=50, extra_tags=" x=1000;"),
        synth.make_email(rng, k, "s.example.com", idx=8, body_len=50, extra_tags=" l=abc;"),
        synth.make_email(rng, k, "s.example.com", idx=10, body_len=50, h=("to", "subject")),
        z.Email(base.from_domain, b"no headers at all", base.public_key),
        z.Email(base.from_domain, b"", base.public_key),
        z.Email(base.from_domain.upper(), base.raw_email, base.public_key),
        synth.mutate(base, "body_flip", rng), synth.mutate(base, "sig_flip", rng), synth.mutate(base, "bh_flip", rng),
    ]
    buf, views = contiguous_views(emails)
    engine.register_host(buf)
    try:
        got = engine.verify_views(views)
    finally:
        engine.unregister_host(buf)
    exp = oracle.verify_batch(emails, now=NOW)
    for i, (g, e) in enumerate(zip(got, exp)):
        assert_records_equal(g, e, i)
    assert [int(g["status"]) for g in got][:5] == [0, 1, 1, 2, 9]


def test_front_end_paths_agree(engine):
    """One mixed batch through the three input paths — staged device front end (pageable views, the default),
    host front end (ZKB_OPT_NO_DEVICE_FRONTEND) and registered memory — gives identical records, equal to the oracle's."""
    from tests.util import contiguous_views
    emails, _ = mixed_emails(seed=17, n_pos=48)
    exp = oracle.verify_batch(emails, now=NOW)
    staged = engine.verify_batch(emails)
    assert engine.last_batch_bytes()["h2d_bytes"] > sum(len(e.raw_email) for e in emails)   # raw messages travelled
    engine.set_flags(z.OPT_NO_DEVICE_FRONTEND)
    try:
        host = engine.verify_batch(emails)
        assert engine.last_batch_bytes()["host_front_end_emails"] == 0     # nothing was declined: nothing was tried
    finally:
        engine.set_flags(0)
    buf, views = contiguous_views(emails)
    engine.register_host(buf)
    try:
        reg = engine.verify_views(views)
    finally:
        engine.unregister_host(buf)
    for i, e in enumerate(exp):
        assert_records_equal(staged[i], e, i)
        assert_records_equal(host[i], e, i)
        assert_records_equal(reg[i], e, i)


def test_differential_fuzz_of_the_front_ends():
    """tools/fuzz_frontend.py (dirty, mutated, multi-signature mail; device front end with host fallback vs oracle),
    a fixed seed of it.  Larger runs: `python tools/fuzz_frontend.py 20000 <seed>` (60 000 mails checked in round 1)."""
    import importlib.util
    import os
    import sys
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fuzz_frontend.py")
    spec = importlib.util.spec_from_file_location("fuzz_frontend", path)
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    argv = sys.argv
    sys.argv = ["fuzz_frontend.py", "3000", "11"]
    try:
        assert fz.main() == 0
    finally:
        sys.argv = argv


def test_concurrent_callers(engine):
    """calls on one engine are serialised, distinct engines are independent — four threads, two
    engines, every result identical to the single-threaded one."""
    import threading
    emails, _ = mixed_emails(seed=41, n_pos=40)
    ref = engine.verify_batch(emails).tobytes()
    other = z.Engine(device=0, now_unix=NOW, host_threads=2)
    errs = []

    def work(eng, rounds):
        try:
            for _ in range(rounds):
                assert eng.verify_batch(emails).tobytes() == ref
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    try:
        ts = [threading.Thread(target=work, args=(engine if i % 2 == 0 else other, 6)) for i in range(4)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    finally:
        other.close()
    assert not errs, errs
