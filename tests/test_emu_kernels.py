"""The kernel SOURCES (sha256.cuh, rsa.cuh, dfa.cuh) compiled for the host through tests/emu and
checked against hashlib / the oracle / Python big ints.  This box has no GPU; the emulation runs
the same source lines (warp shuffles, ballots, PTX carry chains emulated) so that arithmetic and
control-flow bugs are caught before GPU time is spent.  The -m gpu tests remain the authority."""
import hashlib
import math
import re

import numpy as np
import pytest
from cryptography.hazmat.primitives import hashes
from cryptography.hazmat.primitives.asymmetric import padding

import oracle
from tests import emu
from tests.util import key_pool


def test_sha256_kernel_source():
    rng = np.random.default_rng(1)
    lens = [0, 1, 3, 55, 56, 57, 63, 64, 65, 119, 120, 127, 128, 129, 4096, 4097, 520, 1000]
    lens += [int(x) for x in rng.integers(0, 700, size=110)]
    msgs = [rng.integers(0, 256, size=l, dtype=np.uint8).tobytes() for l in lens]
    for prefetch in (False, True):
        got = emu.sha256_batch(msgs, prefetch=prefetch)
        for m, g in zip(msgs, got):
            assert g == hashlib.sha256(m).digest(), (len(m), prefetch)
    for rot in (1, 2, 3):       # rotates issued on the FMA pipe (multiply by a power of two + add of the halves)
        got = emu.sha256_batch(msgs, rot=rot)
        for m, g in zip(msgs, got):
            assert g == hashlib.sha256(m).digest(), (len(m), rot)


def _rsa_cases(bits, n):
    keys = key_pool()[bits]
    ks, ds, ss, exp = [], [], [], []
    for i in range(n):
        k = keys[i % len(keys)]
        m = b"msg%d" % i
        sig = k.private.sign(m, padding.PKCS1v15(), hashes.SHA256())
        d = hashlib.sha256(m).digest()
        kind = i % 6
        if kind == 1:
            sig = bytes([sig[0] ^ 1]) + sig[1:]
        elif kind == 2:
            d = hashlib.sha256(m + b"x").digest()
        elif kind == 3:
            sig = b"\xff" * len(sig)   # s >= n
        elif kind == 4:
            sig = b"\x00" * len(sig)
        ks.append(k.der); ds.append(d); ss.append(sig)
        exp.append(1 if oracle.rsa_verify_sha256(k.der, d, sig) == 1 else 0)
    return ks, ds, ss, exp


@pytest.mark.parametrize("bits,limbs,lanes", [(2048, 64, 4), (2048, 64, 8), (2048, 64, 16), (1024, 32, 2),
                                              (1024, 32, 4), (1024, 32, 8)])
def test_rsa_kernel_source(bits, limbs, lanes):
    ks, ds, ss, exp = _rsa_cases(bits, 12)
    assert emu.rsa_verify(ks, ds, ss, limbs, lanes) == exp
    assert sum(exp) >= 4


def test_rsa_kernel_source_with_dedicated_squaring():
    """rsa_verify_kernel<64, 4, false, SQR = true>: the 16 squarings of s^65537 through Mont::sqr (triangular products
    split over the four lanes, shared-memory combine, fed reduction).  Same verdicts as the oracle, as the plain kernel,
    and on operands built to stress the carry paths (all-ones limbs, tiny values, s = n - 1)."""
    ks, ds, ss, exp = _rsa_cases(2048, 36)
    assert emu.rsa_verify(ks, ds, ss, 64, 104) == exp == emu.rsa_verify(ks, ds, ss, 64, 4)
    assert sum(exp) >= 12
    # 1536- and 2047-bit moduli in the 64-limb class (top limbs zero), signatures that verify and near-misses
    rng = np.random.default_rng(77)
    for nbits in (1536, 2047, 2048):
        while True:
            p, q = _prime(nbits // 2, rng), _prime(nbits - nbits // 2, rng)
            n = p * q
            phi = (p - 1) * (q - 1)
            if n.bit_length() == nbits and math.gcd(65537, phi) == 1:
                break
        d = pow(65537, -1, phi)
        k = (nbits + 7) // 8
        der = _der(n, 65537)
        ks, ds, ss, exp = [], [], [], []
        for i in range(10):
            h = hashlib.sha256(b"sq%d" % i).digest()
            em = b"\x00\x01" + b"\xff" * (k - 54) + b"\x00" + bytes.fromhex("3031300d060960864801650304020105000420") + h
            s = pow(int.from_bytes(em, "big"), d, n)
            if i == 1:
                s ^= 1
            elif i == 2:
                s = n - 1
            elif i == 3:
                s = 1
            elif i == 4:
                s = (1 << (nbits - 1)) - 1      # long runs of one bits
            elif i == 5:
                s = 0
            ks.append(der); ds.append(h); ss.append(s.to_bytes(k, "big"))
            exp.append(1 if oracle.rsa_verify_sha256(der, h, ss[-1]) == 1 else 0)
        assert exp[0] == 1 and exp[1] == 0
        assert emu.rsa_verify(ks, ds, ss, 64, 104) == exp, nbits
        assert emu.rsa_verify(ks, ds, ss, 64, 204) == exp, nbits


def _prime(bits, rng):
    def is_prime(n):
        if n % 2 == 0:
            return False
        d, r = n - 1, 0
        while d % 2 == 0:
            d //= 2; r += 1
        for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
            x = pow(a, d, n)
            if x in (1, n - 1):
                continue
            for _ in range(r - 1):
                x = x * x % n
                if x == n - 1:
                    break
            else:
                return False
        return True
    while True:
        nb = (bits + 7) // 8
        p = (int.from_bytes(rng.bytes(nb), "big") >> (8 * nb - bits)) | (1 << (bits - 1)) | 1
        if all(p % q for q in (3, 5, 7, 11, 13, 17, 19, 23, 29, 31)) and is_prime(p):
            return p


def _der(n, e):
    def ln(l):
        return bytes([l]) if l < 128 else (b"\x81" + bytes([l]) if l < 256 else b"\x82" + l.to_bytes(2, "big"))
    def integer(v):
        b = v.to_bytes((v.bit_length() + 8) // 8, "big")
        return b"\x02" + ln(len(b)) + b
    body = integer(n) + integer(e)
    return b"\x30" + ln(len(body)) + body


def _is_prime(n):
    if n < 2 or n % 2 == 0:
        return n == 2
    if any(n % q == 0 for q in (3, 5, 7, 11, 13, 17, 19, 23, 29, 31)):
        return n in (3, 5, 7, 11, 13, 17, 19, 23, 29, 31)
    d, r = n - 1, 0
    while d % 2 == 0:
        d //= 2; r += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(r - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def _prime_from(x, step):
    x |= 1
    while not (_is_prime(x) and math.gcd(65537, x - 1) == 1):
        x += 2 * step
    return x


def test_rsa_kernel_sources_on_moduli_with_long_carry_runs():
    """Moduli whose limbs are almost all ones or almost all zeros (p, q just below 2^1024; p just above 2^1023 times q just
    above 2^1024; one factor with a long run of zero limbs): the Montgomery steps then add / propagate across whole lanes,
    n0inv takes unusual values and the final conditional subtraction sits at its edge.  Valid signatures must verify in the
    squaring kernel (lanes code 104) and in the plain one, near-misses must not."""
    forms = [
        (_prime_from((1 << 1024) - (1 << 20), -1), _prime_from((1 << 1024) - (1 << 40), -1)),      # n = 0xffff...: leading ones
        (_prime_from((1 << 1023) + (1 << 30), 1), _prime_from((1 << 1024) + (1 << 8), 1)),          # n = 0x8000...0: leading zeros
        (_prime_from((1 << 1023) + (1 << 700) + 12345, 1), _prime_from((3 << 1022) + (1 << 64), 1)),  # zero limbs in the middle
    ]
    for p, q in forms:
        n = p * q
        assert n.bit_length() in (2047, 2048)
        d = pow(65537, -1, (p - 1) * (q - 1))
        k = (n.bit_length() + 7) // 8
        der = _der(n, 65537)
        ks, ds, ss, exp = [], [], [], []
        for i in range(8):
            h = hashlib.sha256(b"carry%d" % i).digest()
            em = b"\x00\x01" + b"\xff" * (k - 54) + b"\x00" + bytes.fromhex("3031300d060960864801650304020105000420") + h
            s = pow(int.from_bytes(em, "big"), d, n)
            if i == 5:
                s ^= 1 << 1500
            elif i == 6:
                s = n - 1
            elif i == 7:
                s = (n + 1) // 2
            ks.append(der); ds.append(h); ss.append(s.to_bytes(k, "big"))
            exp.append(1 if oracle.rsa_verify_sha256(der, h, ss[-1]) == 1 else 0)
        assert exp[:5] == [1] * 5 and exp[5] == 0
        assert emu.rsa_verify(ks, ds, ss, 64, 104) == exp, hex(n)[:20]
        assert emu.rsa_verify(ks, ds, ss, 64, 4) == exp, hex(n)[:20]
        assert emu.rsa_verify(ks, ds, ss, 64, 8) == exp, hex(n)[:20]


@pytest.mark.parametrize("nbits,limbs,lanes,e", [(2047, 64, 8, 65537), (1536, 64, 4, 65537), (2048, 64, 8, 3),
                                                 (1000, 32, 4, 17), (3072, 128, 16, 65537), (4096, 128, 8, 65537)])
def test_rsa_kernel_source_odd_moduli_and_exponents(nbits, limbs, lanes, e):
    """Moduli that do not fill their limb class, short moduli (k < 4*limbs) and e != 65537 (generic
    ladder): hand-made keys, signatures forged with the private exponent."""
    rng = np.random.default_rng(nbits + e)
    while True:
        p, q = _prime(nbits // 2, rng), _prime(nbits - nbits // 2, rng)
        n = p * q
        phi = (p - 1) * (q - 1)
        if n.bit_length() == nbits and math.gcd(e, phi) == 1:
            break
    d = pow(e, -1, phi)
    k = (nbits + 7) // 8
    der = _der(n, e)
    ks, ds, ss, exp = [], [], [], []
    for i in range(6):
        h = hashlib.sha256(b"m%d" % i).digest()
        em = b"\x00\x01" + b"\xff" * (k - 54) + b"\x00" + bytes.fromhex("3031300d060960864801650304020105000420") + h
        s = pow(int.from_bytes(em, "big"), d, n)
        if i == 1:
            s ^= 1 << 7
        if i == 2:
            h = hashlib.sha256(b"other").digest()
        ks.append(der); ds.append(h); ss.append(s.to_bytes(k, "big"))
        exp.append(1 if oracle.rsa_verify_sha256(der, h, ss[-1]) == 1 else 0)
    assert exp == [1, 0, 0, 1, 1, 1]
    assert emu.rsa_verify(ks, ds, ss, limbs, lanes, generic=(e != 65537)) == exp


RUST_PY = [(rb"abc", None), (rb"a+", None), (rb"a*", None), (rb"[a-c]+d", None), (rb"(a|ab)(c|bcd)", None),
           (rb"from:[^\r\n]*@example\.com", None), (rb"subject:[^\r\n]+", None), (rb"Transaction ID: [A-Z0-9]+", None),
           (rb"(?i)hello", None), (rb"a{2,4}", None), (rb"a{2,4}?", None), (rb"a{3}", None), (rb"a{2,}b", None),
           (rb"x*?y", None), (rb"^abc", None), (rb"abc$", None), (rb"(?m)^a+$", None), (rb"\d+\.\d+", None),
           (rb"[^a]+", None), (rb".", None), (rb".*", None), (rb".+?b", None), (rb"(?s).+", None),
           (rb"\w+@\w+\.com", None), (rb"(foo|foobar|fo)", None), (rb"(?:ab)*c", None),
           (rb"[[:alpha:]]+", rb"[A-Za-z]+"), (rb"\x41+", None), (rb"a|b|", None), (rb"\Aab|cd\z", rb"\Aab|cd\Z"),
           (rb"(?i)[k-s]+", None), (rb"a?b?c?", None), (rb"(a*)*b", None), (rb"(a|b)*?c", None), (rb"\s+", None),
           (rb"to:[^\r\n]+\r\n", None), (rb"(?P<n>a)(?<m>b)", rb"(?P<n>a)(?P<m>b)"), (rb"[\]\[]+", None),
           (rb"(?U)a+", rb"a+?"), (rb"a(?i)b|c", rb"a[bB]|[cC]")]


def _py_spans(pat, hay):
    out, last = [], None
    for m in re.finditer(pat, hay):
        s, e = m.span()
        if s == e and last == e:   # Rust skips an empty match adjacent to the previous match
            continue
        out.append((s, e)); last = e
    return out


def _hays(seed):
    rng = np.random.default_rng(seed)
    alpha = b"abcdxy @.:\r\nAB019kK=\x00"
    hays = [b"", b"a", b"abc", b"aaab", b"xabcabcx", b"from:bob <bob@example.com>\r\nsubject:hi there\r\nto:x\r\n",
            b"Transaction ID: A1B2C3 ok", b"hello HELLO HeLLo", b"ab\nabc\naa\n", b"3.14 and 2.71", b"][]["]
    for _ in range(70):
        l = int(rng.integers(0, 40))
        hays.append(bytes(alpha[i] for i in rng.integers(0, len(alpha), size=l)))
    words = [b"from:", b"subject:", b"abc", b"aaa", b" ", b"\r\n", b"Transaction ID: ", b"X9", b"hello", b"=\r\n", b"=", b"@example.com", b"3.14", b"to:"]
    for _ in range(40):   # long haystacks: the unrolled 16-byte block path
        k = int(rng.integers(5, 60))
        hays.append(b"".join(words[i] for i in rng.integers(0, len(words), size=k)))
    return hays


@pytest.mark.parametrize("pat,pypat", RUST_PY)
def test_regex_compiler_oracle_search_and_dfa_kernel_source(pat, pypat):
    """pattern -> ZDF1 (regexc.hpp) -> spans: the oracle's find_iter and the DFA kernel source must
    both equal Python `re` (leftmost-first, with Rust's empty-match iteration rule)."""
    fwd, bwd = emu.regex_compile(pat)
    hays = _hays(7)
    exp = [_py_spans(pypat or pat, h) for h in hays]
    for h, e in zip(hays, exp):
        cnt, spans = oracle.dfa_find_iter(fwd, bwd, h, 64)
        assert cnt == len(e) and spans == e[:64], (pat, h, e, spans)
    for form in (0, 1, 2):   # DIRECT rows (engine's choice for small DFAs), class-compressed, u32 elements
        got = emu.dfa_scan(fwd, bwd, hays, qp=False, use_smem=(form != 2), table_form=form)
        for h, e, r in zip(hays, exp, got):
            c, s, en, panic = (int(x) for x in r)
            assert c == len(e) and not panic and (not e or (s, en) == e[0]), (pat, form, h, e, r)
    # fused quoted-printable soft-break removal vs the oracle on the cleaned copy
    hq = [h.replace(b"b", b"b=\r\n", 1) if i % 2 else h + b"=\r\n" for i, h in enumerate(hays)]
    for form in (0, 1):
        got = emu.dfa_scan(fwd, bwd, hq, qp=True, use_smem=False, table_form=form)
        for h, r in zip(hq, got):
            clean, _ = oracle.qp_clean(h)
            cnt, spans = oracle.dfa_find_iter(fwd, bwd, clean, 4)
            c, s, en, panic = (int(x) for x in r)
            assert c == cnt and not panic and (not cnt or (s, en) == spans[0]), (pat, form, h, clean, spans, r)


def test_regex_unicode_classes_utf8():
    fwd, bwd = emu.regex_compile("é+|[α-ω]+|.".encode())
    hay = "aéé βγ €\n".encode()
    cnt, spans = oracle.dfa_find_iter(fwd, bwd, hay, 64)
    exp = [m.span() for m in re.finditer("é+|[α-ω]+|.".encode().decode(), hay.decode())]
    # convert character spans to byte spans
    s = hay.decode()
    b = lambda i: len(s[:i].encode())
    assert spans == [(b(x), b(y)) for x, y in exp]
    # negated class spans whole scalars, never a lone continuation byte
    fwd, bwd = emu.regex_compile(rb"[^a]")
    assert oracle.dfa_find_iter(fwd, bwd, "a€a".encode(), 8) == (1, [(1, 4)])
    assert oracle.dfa_find_iter(fwd, bwd, b"a\xffa", 8) == (0, [])   # invalid UTF-8 never matches


@pytest.mark.parametrize("bad", [rb"a(", rb"\bfoo", rb"[a", rb"*a", rb"\p{Greek}", rb"(?x)a", rb"a{5,2}", rb"(?-u:.)", rb"a)", rb"[z-a]", rb"\1"])
def test_regex_compiler_rejects(bad):
    with pytest.raises(ValueError):
        emu.regex_compile(bad)


_PIECES = [b"line", b" ", b"\t", b"\r\n", b"\n", b"\r", b"=\r\n", b"  ", b"text text", b"", b" \r\n", b"\r\n\r\n", b"x" * 17, b"\t \t"]


def test_canon_body_kernel_source_vs_oracle():
    """Device-side body canonicalisation (canon.cuh) against the oracle on dirty bodies: tabs, WSP
    runs, SP before CRLF, bare CR/LF, trailing blank lines, missing final CRLF, empty, l= cuts."""
    rng = np.random.default_rng(5)
    bodies = [b"", b"\r\n", b"abc ", b" ", b"\t", b"a \r\n", b" \r\r\n", b"a\r\n\r\n\r\n", b"a \t \r\nb", b"\r", b"\n",
              b" C \r\nD \t E\r\n\r\n\r\n", b"x" * 64, b"y" * 63 + b" ", b"no final newline", b"\r\n\r\n", b" \r\n \r\n"]
    for _ in range(400):
        k = int(rng.integers(0, 14))
        bodies.append(b"".join(_PIECES[i] for i in rng.integers(0, len(_PIECES), size=k)))
    # long bodies of mostly clean text (the 16-byte pass-through path) with dirty spots at every alignment: SP / TAB /
    # CR at block edges, WSP runs and SP CRLF straddling blocks, soft breaks, a dirty byte right after a clean block
    words = [b"alpha", b"be", b"gamma-delta", b"x", b"0123456789", b"Transaction", b"ID:", b"(c)", b"zz;zz", b"_"]
    dirt = [b"  ", b" \t", b"\t", b" \r\n", b"\r\n\r\n", b"\r", b"\n", b" \r", b"\t\r\n", b"=\r\n", b"   \r\n"]
    for i in range(300):
        parts, n = [], int(rng.integers(20, 400))
        line = 0
        for _ in range(n):
            if rng.random() < (0.0 if i % 3 == 0 else 0.04):
                parts.append(dirt[int(rng.integers(0, len(dirt)))])
            wd = words[int(rng.integers(0, len(words)))]
            parts.append(wd)
            line += len(wd) + 1
            if line > 60:
                parts.append(b"\r\n"); line = 0
            else:
                parts.append(b" ")
        body = b"".join(parts)
        if i % 4 == 0:
            body = body.rstrip(b" ") + b"\r\n"
        bodies.append(b"x" * (i % 17) + body)
    for relaxed in (True, False):
        got = emu.canon_bodies(bodies, relaxed=relaxed)
        for b, g in zip(bodies, got):
            assert g == oracle.canon_body(b, relaxed), (relaxed, b, g, oracle.canon_body(b, relaxed))
        for l in (0, 1, 5, 100000):
            got = emu.canon_bodies(bodies[:60], relaxed=relaxed, l=l)
            for b, g in zip(bodies[:60], got):
                assert g == oracle.canon_body(b, relaxed)[:l], (relaxed, l, b, g)


def _fe_compare(raw: bytes, dom: bytes, k=256, limbs=64, allow_skip=False, same=True) -> int:
    """Both device front ends on one message against the host front end: the scalar twin (frontend.cuh) and the
    warp-cooperative kernel source (frontend_warp.cuh, 32 emulated lanes).  Returns the scalar verdict (0 declined,
    1 accepted and byte-identical to the host, 2 mail parse error on both sides); the warp form must never mismatch,
    may decline more, and with same=True must reach the same verdict."""
    import ctypes as C
    sc = C.c_int(-99)
    rw = emu.lib().emu_fe_compare_warp(raw, len(raw), dom, len(dom), k, limbs, 1 if allow_skip else 0, C.byref(sc))
    r = sc.value
    assert rw >= 0, ("warp front end differs from the host", rw, raw)
    if r >= 0:
        assert rw == r or (rw == 0 and not same), ("warp / scalar verdicts", rw, r, raw)
    return r


def test_device_front_end_source_on_synthetic_mail():
    """frontend.cuh (device-side header split / tag list / header selection / preimage / base64) against
    the host front end: live results must be byte-identical; well-formed mail must not fall back."""
    from tests.util import mixed_emails
    for seed, tok in ((21, False), (22, True)):
        emails, labels = mixed_emails(seed=seed, with_token=tok)
        for e, lab in zip(emails, labels):
            big = len(e.public_key.key) > 200
            r = _fe_compare(e.raw_email, e.from_domain.encode(), 256 if big else 128, 64 if big else 32)
            assert r >= 0, (lab, r)
            if lab in ("pos", "body_flip", "sig_flip", "wrong_key", "bh_flip", "header_flip"):
                assert r == 1, (lab, r)
            if lab in ("domain_mismatch", "missing_tag"):
                assert r == 0, (lab, r)


from hypothesis import given, settings, strategies as st  # noqa: E402

_FE_NAME = st.sampled_from([b"From", b"from", b"FROM", b"To", b"Subject", b"Date", b"Message-ID", b"X-Test", b"Cc", b"Received", b"Subject "])
_FE_VAL = st.lists(st.sampled_from([b"a", b"B", b" ", b"\t", b"  ", b"\r\n ", b"\r\n\t", b"x@y.z", b";", b"=", b"\xc3\xa9", b"\r", b"\n "]), max_size=8).map(b"".join)
_FE_SEP = st.sampled_from([b":", b": ", b":  ", b":\t", b" :"])
_FE_BODY = st.lists(st.sampled_from([b"line", b" ", b"\r\n", b"\n", b"=\r\n", b"text"]), max_size=6).map(b"".join)
_FE_B = st.sampled_from([b"QUJD\r\n\t REVG", b"QUJDREVG", b"QUJDRA==", b"QUJ", b"QU JD", b"", b"=QUJD", b"QUJD;", b"QUJD; z=1", b"QUJD ;\r\n"])


@settings(max_examples=400, deadline=None)
@given(st.lists(st.tuples(_FE_NAME, _FE_SEP, _FE_VAL), min_size=1, max_size=7), _FE_BODY,
       st.sampled_from(["relaxed/relaxed", "simple/simple", "relaxed/simple", "simple/relaxed", "relaxed", "simple", "bogus", None]),
       st.lists(st.sampled_from(["from", "to", "subject", "date", "cc", "x-test", "From", "message-id", "missing", "subject "]), min_size=0, max_size=6),
       st.sampled_from(["", " l=5;", " i=@example.com;", " x=99999999999;", " q=dns/txt;", " z=1;", " v=2;", " d=other.org;", " a=rsa-sha1;",
                        " i=user@sub.example.com; q=dns/txt; x=12345;", " i=@elsewhere.org;", " i=@Example.com;", " q=other;", " x=-5;", " x=1\r\n\t2;",
                        " x=;", " i=;", " x=99999999999999999999;", " i=a@ex\r\n ample.com;"]),
       _FE_B, st.sampled_from(["top", "bottom", "both", "foreign", "foreign", "foreign_sha1", "foreign_after"]),
       st.sampled_from([b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA=", b"AAAA", b"AAAAAAAAAAAAAAAAAAAAAA\r\n\tAAAAAAAAAAAAAAAAAAAAA="]),
       st.booleans())
def test_device_front_end_source_on_dirty_mail(headers, body, canon, hnames, extra, bval, where, bh, allow_skip):
    """Whatever the device front end accepts must equal the host front end byte for byte; anything it
    does not accept must be flagged for the host (never a silent difference)."""
    block = b"".join(k + s + v.rstrip(b"\r\n\t ") + b"\r\n" if not v.endswith((b"\r\n ", b"\r\n\t", b"\n ")) else k + s + v + b"x\r\n"
                     for k, s, v in headers)
    h = ":".join(["from"] + hnames)
    ctag = f" c={canon};" if canon else ""
    sig = (f"DKIM-Signature: v=1; a=rsa-sha256;{ctag} d=example.com; s=s;\r\n\th={h};{extra}\r\n\tbh=").encode() + bh + b";\r\n\tb=" + bval + b"\r\n"
    foreign = b"DKIM-Signature: v=1; a=rsa-sha256; c=relaxed/relaxed; d=Other.org; s=x;\r\n\th=from:to;\r\n\tbh=" + bh + b";\r\n\tb=QUJD\r\n"
    if where == "top":
        raw = sig + block
    elif where == "bottom":
        raw = block + sig
    elif where == "foreign":          # a signature of another domain first: skipped by the reference
        raw = foreign + block + sig
    elif where == "foreign_sha1":     # one the device cannot judge: must fall back
        raw = foreign.replace(b"rsa-sha256", b"rsa-sha1") + sig + block
    elif where == "foreign_after":
        raw = sig + block + foreign
    else:
        raw = sig + block + sig
    raw += b"\r\n" + body
    r = _fe_compare(raw, b"Example.COM", k=6, limbs=32, allow_skip=allow_skip, same=False)
    assert r >= 0, (r, raw)
    if where == "foreign_sha1" or (where == "foreign" and not allow_skip):
        assert r in (0, 2), (r, raw)


def test_device_front_end_accepts_passing_optional_tags():
    from zkemail_rs_b200 import synth
    rng = np.random.default_rng(8)
    k = key_pool()[2048][0]
    cases = [(" i=@mail.example.com; q=dns/txt; x=9999999999;", 1), (" i=user@sub.mail.example.com;", 1), (" t=1700000000; x=1700000\r\n\t900;", 1),
             (" i=@other.example.com;", 0), (" q=dns;", 0), (" x=0;", 1), (" x=abc;", 0), (" l=10;", 1), (" l=+7;", 1), (" l=0;", 1),
             (" l=99999999999;", 1), (" l=;", 0), (" l=1a;", 0), (" l=-1;", 0)]
    for extra, want in cases:
        e = synth.make_email(rng, k, "mail.example.com", idx=3, body_len=200, extra_tags=extra)
        assert _fe_compare(e.raw_email, b"mail.example.com") == want, extra


def test_device_front_end_skips_foreign_signatures_only_when_allowed():
    from tests.util import mixed_emails
    from zkemail_rs_b200 import synth
    emails, labels = mixed_emails(seed=23)
    good = [e for e, lab in zip(emails, labels) if lab == "pos"][:6]
    other = synth.make_email(np.random.default_rng(4), key_pool()[2048][1], "elsewhere.example.org", idx=9, body_len=80).raw_email
    foreign = other[: other.find(b"\r\n", other.find(b"\tb=")) + 2]
    assert foreign.startswith(b"DKIM-Signature")
    for e in good:
        big = len(e.public_key.key) > 200
        args = (e.from_domain.encode(), 256 if big else 128, 64 if big else 32)
        top = e.raw_email.startswith(b"DKIM-Signature")
        assert _fe_compare(foreign + e.raw_email, *args, allow_skip=True) == 1
        assert _fe_compare(foreign + e.raw_email, *args, allow_skip=False) == 0
        if top:   # candidate first, foreign signature later: fine with regex parts as well
            cut = e.raw_email.find(b"\r\n", e.raw_email.find(b"\tb=")) + 2
            assert _fe_compare(e.raw_email[:cut] + foreign + e.raw_email[cut:], *args, allow_skip=False) == 1


def test_regex_unicode_perl_classes():
    r"""\d \w \s follow the Unicode definitions in (?u) mode (as DFARegex::new compiles them) and the ASCII
    ones under (?-u)."""
    def spans(pat, hay):
        fwd, bwd = emu.regex_compile(pat.encode())
        return oracle.dfa_find_iter(fwd, bwd, hay.encode(), 64)[1]
    def bytespans(pat, hay, flags=0):
        b = lambda i: len(hay[:i].encode())
        return [(b(m.start()), b(m.end())) for m in re.finditer(pat, hay, flags)]
    hay = "id ١٢٣ and 456, héllo wörld_ok\u00a0x\u2003y naïve"
    assert spans(r"\d+", hay) == bytespans(r"\d+", hay)
    assert spans(r"\w+", hay) == bytespans(r"\w+", hay)
    assert spans(r"\s+", hay) == bytespans(r"\s+", hay)
    assert spans(r"\D+", "12ab٣") == bytespans(r"\D+", "12ab٣")
    assert spans(r"(?-u:\d+)", hay) == bytespans(r"\d+", hay, re.ASCII)
    assert spans(r"(?-u:\w+)", "héllo") == [(0, 1), (3, 6)]
    assert spans(r"[\d_]+", "a_١_9") == bytespans(r"[\d_]+", "a_١_9")
