"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI of
libzkemail_b200.so; the CPU oracle (oracle/) is only the checker."""
import hashlib

import numpy as np
import pytest

import oracle
import zkemail_rs_b200 as z
from zkemail_rs_b200 import synth
from zkemail_rs_b200.structs import CompiledRegex, RegexInfo, RegexPattern
from tests.util import NOW, assert_records_equal, key_pool, mixed_emails

pytestmark = pytest.mark.gpu


def test_sha256_kernel_vs_hashlib_and_oracle(engine):
    rng = np.random.default_rng(3)
    lens = [0, 1, 3, 55, 56, 57, 63, 64, 65, 119, 120, 127, 128, 129, 4096, 4097, 520, 1000, 102400]
    lens += [int(x) for x in rng.integers(0, 3000, size=500)]
    msgs = [rng.integers(0, 256, size=l, dtype=np.uint8).tobytes() for l in lens]
    msgs[1] = b"a"
    msgs.append(b"abc")  # FIPS 180-4 KAT
    got = engine.sha256_batch(msgs)
    assert got[-1].hex() == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"
    for m, g in zip(msgs, got):
        assert g == hashlib.sha256(m).digest()
        assert g == oracle.sha256(m)


@pytest.mark.parametrize("lanes", [2, 4, 8, 16])
def test_rsa_kernel_vs_oracle(lanes):
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.asymmetric import padding
    eng = z.Engine(rsa_lanes=lanes)
    try:
        rng = np.random.default_rng(5)
        keys = key_pool()
        ks, ds, ss = [], [], []
        for i in range(96):
            bits = 1024 if i % 3 == 2 else 2048
            k = keys[bits][i % len(keys[bits])]
            m = b"message %d" % i
            sig = k.private.sign(m, padding.PKCS1v15(), hashes.SHA256())
            d = hashlib.sha256(m).digest()
            kind = i % 8
            if kind == 1:
                sig = bytes([sig[0] ^ 0x01]) + sig[1:]
            elif kind == 2:
                d = hashlib.sha256(m + b"!").digest()
            elif kind == 3:
                sig = sig[1:]                      # wrong length
            elif kind == 4:
                sig = b"\xff" * len(sig)           # s >= n
            elif kind == 5:
                sig = b"\x00" * len(sig)           # s = 0
            elif kind == 6:
                k = keys[bits][(i + 1) % len(keys[bits])]  # wrong key
            ks.append(k.der); ds.append(d); ss.append(sig)
        got = eng.rsa_verify_batch(ks, ds, ss)
        exp = [oracle.rsa_verify_sha256(k, d, s) for k, d, s in zip(ks, ds, ss)]
        assert got == exp
        assert sum(got) >= 24
        # key rejected by the loader
        assert eng.rsa_verify_batch([b"\x30\x03\x02\x01\x01"], [b"\0" * 32], [b"\1"]) == [2]
    finally:
        eng.close()


PATTERNS = [r"abc", r"a+", r"a*", r"(a|ab)(c|bcd)", r"from:[^\r\n]*@example\.com", r"subject:[^\r\n]+",
            r"Transaction ID: [A-Z0-9]+", r"(?i)hello", r"a{2,4}", r"^abc", r"abc$", r"(?m)^a+$", r".*",
            r"[^a]+", r"(foo|foobar|fo)", r"\d+\.\d+", r"to:[^\r\n]+\r\n", r"\w+@\w+\.com"]


@pytest.mark.parametrize("qp", [False, True])
def test_dfa_kernel_vs_oracle(engine, qp):
    rng = np.random.default_rng(11)
    alpha = b"abcdxy @.:\r\nAB019kK=\x00"
    hays = [b"", b"a", b"abc", b"aaab", b"xabcabcx", b"from:bob <bob@example.com>\r\nsubject:hi there\r\nto:x\r\n",
            b"Transaction ID: A1B2C3 ok", b"hello HELLO", b"ab\nabc\naa\n", b"3.14 and 2.71", b"=\r\n", b"a=\r\nbc=\r\n"]
    for _ in range(300):
        l = int(rng.integers(0, 200))
        hays.append(bytes(alpha[i] for i in rng.integers(0, len(alpha), size=l)))
    if qp:
        hays = [h.replace(b"b", b"b=\r\n", 1) if i % 2 else h + b"=\r\n" for i, h in enumerate(hays)]
    for pat in PATTERNS:
        dfa = z.compile_regex(pat)
        got = engine.dfa_scan_batch(dfa, hays, qp=qp)
        for h, r in zip(hays, got):
            hay = oracle.qp_clean(h)[0] if qp else h
            cnt, spans = oracle.dfa_find_iter(dfa.fwd, dfa.bwd, hay, 4)
            assert int(r[0]) == cnt, (pat, h, r, cnt, spans)
            if cnt:
                assert (int(r[1]), int(r[2])) == spans[0], (pat, h, r, spans)


def test_verify_batch_vs_oracle(engine):
    emails, labels = mixed_emails(seed=21)
    got = engine.verify_batch(emails)
    exp = oracle.verify_batch(emails, now=NOW)
    assert len(got) == len(exp)
    for g, e, lab in zip(got, exp, labels):
        assert_records_equal(g, e, lab)
    st = [int(g["status"]) for g in got]
    assert st.count(0) == labels.count("pos")
    assert all(s != 0 for s, lab in zip(st, labels) if lab != "pos")


def test_verify_with_regex_vs_oracle(engine):
    emails, labels = mixed_emails(seed=22, with_token=True)
    # compile against the first email's haystacks (helpers/src/generator.rs:63-78)
    hdr, body = oracle.canonicalize_signed_email(emails[0].raw_email, NOW)
    clean, _ = oracle.qp_clean(body)
    header_parts = z.compile_regex_parts([RegexPattern(r"from:[^\r\n]*@mail[0-9]\.example\.com", None),
                                          RegexPattern(r"\r\nsubject:([^\r\n]+)\r\n", None)], hdr)
    body_parts = z.compile_regex_parts([RegexPattern(r"Transaction ID: ([A-Z0-9]+)", None)], clean)
    for p in header_parts + body_parts:
        p.captures = None  # per-email captures differ; checked separately below
    info = RegexInfo(header_parts, body_parts)
    got = engine.verify_with_regex_batch(emails, info)
    exp = oracle.verify_batch(emails, header_parts, body_parts, now=NOW)
    for g, e, lab in zip(got, exp, labels):
        assert_records_equal(g, e, lab)
    assert sum(int(g["status"]) == 0 for g in got) == labels.count("pos")
    # captures: correct for email 0, wrong for the others -> identical verdicts on both sides
    body_parts2 = z.compile_regex_parts([RegexPattern(r"Transaction ID: ([A-Z0-9]+)", [1])], clean)
    info2 = RegexInfo(None, body_parts2)
    got2 = engine.verify_with_regex_batch(emails[:6], info2)
    exp2 = oracle.verify_batch(emails[:6], None, body_parts2, now=NOW)
    for g, e in zip(got2, exp2):
        assert_records_equal(g, e, "captures")
    assert int(got2[0]["status"]) == 0 and int(got2[1]["status"]) == 7


def test_single_email_api(engine):
    emails, labels = mixed_emails(seed=23, n_pos=3)
    out = engine.verify_email(emails[0])
    assert out.from_domain_hash == hashlib.sha256(emails[0].from_domain.encode()).digest()
    assert out.public_key_hash == hashlib.sha256(emails[0].public_key.key).digest()
    with pytest.raises(z.VerificationPanic) as ei:
        engine.verify_email(emails[3])  # body_flip
    assert ei.value.status == 3


def test_resident_batch_matches_pipeline(engine):
    emails, labels = mixed_emails(seed=24)
    from zkemail_rs_b200.engine import EmailViews
    views = EmailViews.from_emails(emails)
    a = engine.verify_views(views)
    pb = engine.prepare(views)
    pb.run(); pb.run()
    b = pb.fetch()
    st = pb.stats()
    pb.close()
    assert a.tobytes() == b.tobytes()
    assert st["n_emails"] == len(emails) and st["kernel_launches"] >= 3


@pytest.mark.parametrize("flags", [0, z.OPT_NO_OVERLAP])
def test_resident_multi_chunk_overlapped_schedule(flags):
    """zkb_batch_run_async with several resident chunks: hashing of chunk k+1 on the side stream next to the RSA of
    chunk k (and the single-stream order) gives the records of the plain pipeline and of the oracle."""
    from zkemail_rs_b200.engine import EmailViews
    eng = z.Engine(device=0, now_unix=NOW, chunk_emails=8, flags=flags)     # resident chunks hold 32 emails
    try:
        emails, _ = mixed_emails(seed=31, n_pos=150, with_token=True)
        info = RegexInfo(None, [CompiledRegex(z.compile_regex(r"Transaction ID: [A-Z0-9]+"), None)])
        views = EmailViews.from_emails(emails)
        rs = z.RegexSet(eng, info)
        pb = eng.prepare(views, rs, with_captures=False)
        assert len(pb.device_flags()) >= 4
        for _ in range(3):
            pb.run_async()
        got = pb.fetch()
        pb.close()
        exp = oracle.verify_batch(emails, None, info.body_parts, now=NOW)
        for i, (g, e) in enumerate(zip(got, exp)):
            assert_records_equal(g, e, i)
    finally:
        eng.close()


def test_direct_mode_device_canonicalisation(engine):
    """Zero-copy path: raw messages in registered host memory; headers parsed / preimages built / base64
    decoded by frontend.cuh and bodies canonicalised by canon.cuh on the device (irregular messages fall
    back to the host front end).  Must equal the oracle AND the two host-side paths record for record."""
    from tests.util import contiguous_views
    emails, labels = mixed_emails(seed=31, with_token=True)
    rng = np.random.default_rng(32)
    k = key_pool()[2048][0]
    dirty = [b"Trailing space \r\nTab\there  \r\n\r\n\r\n", b"no final newline", b"", b"\r\n", b" \r\n \r\n", b"a\tb  c \r\n\r\n",
             b"bare\nlf and\rcr\r\n", b"x" * 1000 + b" \r\n" + b"y" * 70 + b"\r\n\r\n\r\n\r\n", b"abc "]
    for i, body in enumerate(dirty):
        for canon in ("relaxed/relaxed", "simple/simple"):
            raw = synth.sign_email(synth.default_headers(rng, "mail0.example.com", 200 + i), body, k, "mail0.example.com", canon=canon)
            emails.append(z.Email("mail0.example.com", raw, z.PublicKey(k.der, "rsa")))
            labels.append("pos")
    buf, views = contiguous_views(emails)
    exp = oracle.verify_batch(emails, now=NOW)
    engine.register_host(buf)
    try:
        got = engine.verify_views(views)
        pb = engine.prepare(views)
        pb.run()
        got_res = pb.fetch()
        st = pb.stats()
        pb.close()
        engine.set_flags(z.OPT_NO_DEVICE_FRONTEND)      # registered memory, host front end + device canonicalisation
        mid = engine.verify_views(views)
        engine.set_flags(z.OPT_NO_DIRECT | z.OPT_NO_DEVICE_FRONTEND)   # everything on the host threads
        host = engine.verify_views(views)
    finally:
        engine.set_flags(0)
        engine.unregister_host(buf)
    assert got.tobytes() == mid.tobytes()
    for g, e, lab in zip(got, exp, labels):
        assert_records_equal(g, e, lab)
    assert got.tobytes() == host.tobytes() == got_res.tobytes()
    # ("abc " without a final CRLF is cfdkim's documented quirk: the RFC signer strips the SP, cfdkim keeps it)
    assert sum(int(g["status"]) == 0 for g in got) >= labels.count("pos") - 2
    # with regex + captures: the haystack is the device-canonicalised body; captures are checked on the host
    hdr, body = oracle.canonicalize_signed_email(emails[0].raw_email, NOW)
    clean, _ = oracle.qp_clean(body)
    body_parts = z.compile_regex_parts([RegexPattern(r"Transaction ID: ([A-Z0-9]+)", [1])], clean)
    info = RegexInfo(None, body_parts)
    from zkemail_rs_b200.engine import RegexSet
    rs = RegexSet(engine, info)
    engine.register_host(buf)
    try:
        got2 = engine.verify_views(views, rs)
    finally:
        engine.unregister_host(buf)
        rs.close()
    exp2 = oracle.verify_batch(emails, None, body_parts, now=NOW)
    for g, e, lab in zip(got2, exp2, labels):
        assert_records_equal(g, e, lab)
    assert int(got2[0]["status"]) == 0


def test_raw_resident_batches_redo_the_front_end_every_run(engine):
    """zkb_batch_prepare_raw: raw messages resident in HBM (staged copies of pageable mail, or the DMA'd registered span);
    every run starts at the device front end and ends with the device-built result records; messages the device
    declines are finished by the host front end in fetch.  Records equal the pipeline's and the oracle's."""
    from tests.util import contiguous_views
    from zkemail_rs_b200.engine import EmailViews
    emails, labels = mixed_emails(seed=41, n_pos=60, with_token=True)
    info = RegexInfo([CompiledRegex(z.compile_regex(r"\r\nsubject:[^\r\n]+\r\n"), None)],
                     [CompiledRegex(z.compile_regex(r"Transaction ID: [A-Z0-9]+"), ["Transaction ID"])])
    exp = oracle.verify_batch(emails, info.header_parts, info.body_parts, now=NOW)
    rs = z.RegexSet(engine, info)
    try:
        for registered in (False, True):
            if registered:
                buf, views = contiguous_views(emails)
                engine.register_host(buf)
            else:
                views = EmailViews.from_emails(emails)
            try:
                for with_caps in (False, True):
                    pb = engine.prepare(views, rs, with_captures=with_caps, raw=True)
                    pb.run(); pb.run_async(); pb.run()
                    got = pb.fetch()
                    t = pb.timing_ms()
                    st = pb.stats()
                    pb.close()
                    assert t["front_end_canon"] > 0 and st["kernel_launches"] >= 7
                    pipe = engine.verify_views(views, rs, with_captures=with_caps)
                    assert got.tobytes() == pipe.tobytes()
                    if with_caps:
                        for g, e, lab in zip(got, exp, labels):
                            assert_records_equal(g, e, lab)
            finally:
                if registered:
                    engine.unregister_host(buf)
    finally:
        rs.close()


def test_rsa_kernel_variant_with_dedicated_squaring():
    """rsa_verify_kernel<64, 4, false, SQR> (Mont::sqr for the 16 squarings of s^65537, the default) against the plain
    kernel (ZKB_OPT_NO_SQR).  Same verdicts as the plain kernel and the oracle on forged / flipped / out-of-range signatures, and on a mixed batch end to end."""
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.asymmetric import padding
    eng = z.Engine(now_unix=NOW)
    ref = z.Engine(flags=z.OPT_NO_SQR, now_unix=NOW)
    try:
        keys = key_pool()[2048]
        ks, ds, ss = [], [], []
        for i in range(300):
            k = keys[i % len(keys)]
            m = b"sqr%d" % i
            sig = k.private.sign(m, padding.PKCS1v15(), hashes.SHA256())
            d = hashlib.sha256(m).digest()
            if i % 5 == 1:
                sig = bytes([sig[0] ^ 1]) + sig[1:]
            elif i % 5 == 2:
                d = hashlib.sha256(m + b"x").digest()
            elif i % 5 == 3:
                sig = (b"\xff" * len(sig)) if i % 2 else (b"\x00" * (len(sig) - 1) + b"\x01")
            ks.append(k.der); ds.append(d); ss.append(sig)
        got = eng.rsa_verify_batch(ks, ds, ss)
        assert got == ref.rsa_verify_batch(ks, ds, ss)
        assert got == [1 if oracle.rsa_verify_sha256(k, d, s) == 1 else 0 for k, d, s in zip(ks, ds, ss)]
        assert sum(got) >= 100
        emails, labels = mixed_emails(seed=61, n_pos=40)
        a, b = eng.verify_batch(emails), ref.verify_batch(emails)
        assert a.tobytes() == b.tobytes()
        for g, e, lab in zip(a, oracle.verify_batch(emails, now=NOW), labels):
            assert_records_equal(g, e, lab)
    finally:
        eng.close(); ref.close()


def test_rsa_kernels_on_moduli_with_long_carry_runs():
    """The special-form moduli of tests/test_emu_kernels.py (limbs almost all ones / zeros) through the real kernels: the
    default one (dedicated squaring) and the plain one must both agree with the oracle."""
    from tests.test_emu_kernels import _der, _prime_from
    forms = [
        (_prime_from((1 << 1024) - (1 << 20), -1), _prime_from((1 << 1024) - (1 << 40), -1)),
        (_prime_from((1 << 1023) + (1 << 30), 1), _prime_from((1 << 1024) + (1 << 8), 1)),
        (_prime_from((1 << 1023) + (1 << 700) + 12345, 1), _prime_from((3 << 1022) + (1 << 64), 1)),
    ]
    ks, ds, ss = [], [], []
    for p, q in forms:
        n = p * q
        d = pow(65537, -1, (p - 1) * (q - 1))
        k = (n.bit_length() + 7) // 8
        der = _der(n, 65537)
        for i in range(40):
            h = hashlib.sha256(b"carry%d" % i).digest()
            em = b"\x00\x01" + b"\xff" * (k - 54) + b"\x00" + bytes.fromhex("3031300d060960864801650304020105000420") + h
            s_ = pow(int.from_bytes(em, "big"), d, n)
            if i % 8 == 5:
                s_ ^= 1 << (37 * i % 2000)
            elif i % 8 == 6:
                s_ = n - 1
            elif i % 8 == 7:
                s_ = (n + 1) // 2
            ks.append(der); ds.append(h); ss.append(s_.to_bytes(k, "big"))
    exp = [1 if oracle.rsa_verify_sha256(k, d, s_) == 1 else 0 for k, d, s_ in zip(ks, ds, ss)]
    assert sum(exp) == 3 * 25
    eng, ref = z.Engine(now_unix=NOW), z.Engine(flags=z.OPT_NO_SQR, now_unix=NOW)
    try:
        assert eng.rsa_verify_batch(ks, ds, ss) == exp
        assert ref.rsa_verify_batch(ks, ds, ss) == exp
    finally:
        eng.close(); ref.close()
