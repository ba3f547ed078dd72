"""The product's host front end (dkim_host.hpp, reached through the C ABI without a device) against
the oracle: header split, tag parsing, header selection and canonicalisation are two independent
implementations of the same cfdkim/mailparse behaviour (single pass vs the reference's multi pass).
Also: the C-ABI library loads on a CPU-only box, exports every symbol include/zkemail_b200.h
declares, and refuses to create an engine without a device (no CPU fallback)."""
import os
import re

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle
import zkemail_rs_b200 as z
from zkemail_rs_b200 import engine as zeng, synth
from tests.util import NOW, ROOT, key_pool, mixed_emails


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "zkemail_b200.h")).read()
    declared = set(re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    lib = zeng.load_library()
    for sym in sorted(declared):
        assert getattr(lib, sym) is not None, sym
    assert declared == set(zeng.EXPORTED_SYMBOLS)
    assert lib.zkb_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(z.EngineUnavailable):
        z.Engine()


def _canon_both(raw):
    o = oracle.canonicalize_signed_email(raw, NOW)
    try:
        g = z.canonicalize_signed_email(raw, NOW)
    except z.VerificationPanic:
        g = None
    return o, g


def test_canonicalize_matches_oracle_on_synthetic_mail():
    for seed in (1, 2):
        emails, _ = mixed_emails(seed=seed, with_token=(seed == 2))
        for e in emails:
            o, g = _canon_both(e.raw_email)
            assert o == g


_NAME = st.sampled_from([b"From", b"from", b"FROM", b"To", b"Subject", b"Date", b"Message-ID", b"X-Test", b"Cc", b"Received"])
_VAL = st.lists(st.sampled_from([b"a", b"B", b" ", b"\t", b"  ", b"\r\n ", b"\r\n\t", b"x@y.z", b";", b"=", b"\xc3\xa9", b"\xff"]), max_size=8).map(b"".join)
_SEP = st.sampled_from([b":", b": ", b":  ", b":\t", b" :"])
_BODY = st.lists(st.sampled_from([b"line", b" ", b"\t", b"\r\n", b"\n", b"\r", b"=\r\n", b"  ", b"text text", b""]), max_size=14).map(b"".join)


@settings(max_examples=250, deadline=None)
@given(st.lists(st.tuples(_NAME, _SEP, _VAL), min_size=1, max_size=7), _BODY,
       st.sampled_from(["relaxed/relaxed", "simple/simple", "relaxed/simple", "simple/relaxed"]),
       st.lists(st.sampled_from(["from", "to", "subject", "date", "cc", "x-test", "From", "message-id", "missing"]), min_size=0, max_size=6),
       st.sampled_from(["", " l=5;", " l=0;", " l=99999;", " i=@example.com;", " x=99999999999;"]))
def test_canonicalize_matches_oracle_on_dirty_mail(headers, body, canon, hnames, extra):
    """Arbitrary (often malformed) header blocks: both sides must agree on the preimage bytes, on the
    canonical body, and on when the reference would panic."""
    block = b"".join(k + s + v.rstrip(b"\r\n\t ") + b"\r\n" if not v.endswith((b"\r\n ", b"\r\n\t")) else k + s + v + b"x\r\n"
                     for k, s, v in headers)
    h = ":".join(["from"] + hnames)
    sig = (f"DKIM-Signature: v=1; a=rsa-sha256; c={canon}; d=example.com; s=s;\r\n\th={h};{extra}\r\n"
           f"\tbh=AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA=;\r\n\tb=QUJD\r\n\t REVG\r\n").encode()
    raw = sig + block + b"\r\n" + body
    o, g = _canon_both(raw)
    assert o == g, (raw, o, g)
    raw2 = block + sig + b"\r\n" + body
    o, g = _canon_both(raw2)
    assert o == g, (raw2, o, g)


def test_regex_compile_entry_point():
    d = z.compile_regex(r"subject:[^\r\n]+")
    assert d.fwd[:4] == b"ZDF1" and d.bwd[:4] == b"ZDF1"
    assert oracle.dfa_find_iter(d.fwd, d.bwd, b"to:x\r\nsubject:hello\r\n") == (1, [(6, 19)])
    with pytest.raises(z.RegexError):
        z.compile_regex(r"\bword")
    with pytest.raises(z.RegexError):
        z.compile_regex(r"(unclosed")


def test_compile_regex_parts_mirror():
    from zkemail_rs_b200.structs import RegexPattern
    hay = b"from:Bob <bob@example.com>\r\nsubject:Order 42\r\n"
    parts = z.compile_regex_parts([RegexPattern(r"subject:Order ([0-9]+)", [1])], hay)
    assert parts[0].captures == ["42"]
    with pytest.raises(z.RegexError):
        z.compile_regex_parts([RegexPattern(r"o", None)], hay)      # more than one match
    with pytest.raises(z.RegexError):
        z.compile_regex_parts([RegexPattern(r"zzz", None)], hay)    # no match
