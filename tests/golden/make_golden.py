"""Generates tests/golden/*.json — committed fixtures for the hot path.

The reference (zkemail/zkemail.rs) ships no golden vectors and cannot be run here, so the expected
values are computed by implementations INDEPENDENT of the oracle and of the engine:
  * canonical body / header preimage: zkemail.rs_b200/synth.py (written from RFC 6376 §3.4),
  * hashes: hashlib,  * RSA verdicts: `cryptography` (OpenSSL),  * regex spans: Python `re`.
Run from the repo root:  python tests/golden/make_golden.py
"""
import base64
import hashlib
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from zkemail_rs_b200 import synth  # noqa: E402
from tests.util import key_pool  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
b64 = lambda b: base64.b64encode(b).decode()


def emails():
    rng = np.random.default_rng(20241018)
    keys = key_pool()
    out = []
    canons = ["relaxed/relaxed", "simple/simple", "relaxed/simple", "simple/relaxed"]
    sizes = [0, 1, 64, 300, 4096, 4097]
    for i in range(12):
        bits = 1024 if i % 4 == 3 else 2048
        k = keys[bits][i % len(keys[bits])]
        dom = f"golden{i % 2}.example.com"
        body = synth.synth_body(rng, sizes[i % len(sizes)]) if sizes[i % len(sizes)] else b""
        if i == 5:
            body = b"Trailing space \r\nTab\there  \r\n\r\n\r\n"
        if i == 7:
            body = b"no final newline"
        meta = {}
        raw = synth.sign_email(synth.default_headers(rng, dom, i), body, k, dom, canon=canons[i % 4],
                               sig_position="top" if i % 2 == 0 else "bottom", meta=meta)
        out.append(dict(
            name=f"pos{i}", from_domain=dom, raw_email=b64(raw), key=b64(k.der), key_type="rsa",
            expect=dict(status=0, dkim_detail=0, bh_ok=1, rsa_ok=1,
                        body_hash=hashlib.sha256(meta["canonical_body"]).hexdigest(),
                        header_hash=hashlib.sha256(meta["header_preimage"]).hexdigest(),
                        from_domain_hash=hashlib.sha256(dom.encode()).hexdigest(),
                        public_key_hash=hashlib.sha256(k.der).hexdigest()),
            canonical_body=b64(meta["canonical_body"]), header_preimage=b64(meta["header_preimage"])))
    base = synth.make_email(rng, keys[2048][0], "golden0.example.com", idx=99, body_len=600)
    detail = dict(body_flip=11, sig_flip=13, wrong_key=13, bh_flip=11, header_flip=13, domain_mismatch=1, missing_tag=3)
    for kind in synth.NEGATIVE_KINDS:
        e = synth.mutate(base, kind, rng, other_key=keys[2048][1])
        out.append(dict(name=kind, from_domain=e.from_domain, raw_email=b64(e.raw_email), key=b64(e.public_key.key),
                        key_type="rsa", expect=dict(status=3, dkim_detail=detail[kind])))
    return out


def py_spans(pat, hay):
    out, last = [], None
    for m in re.finditer(pat, hay):
        s, e = m.span()
        if s == e and last == e:
            continue
        out.append([s, e]); last = e
    return out


def regexes():
    pats = [r"from:[^\r\n]*@example\.com", r"\r\nsubject:[^\r\n]+\r\n", r"Transaction ID: [A-Z0-9]+", r"a*", r"(a|ab)(c|bcd)",
            r"(?i)order [0-9]{2,4}", r"(?m)^to:.*$", r"[^ ]+@[a-z.]+"]
    hays = [b"from:Bob <bob@example.com>\r\nto:al@example.org\r\nsubject:Order 42 shipped\r\n",
            b"Your Transaction ID: ZX81AB and Transaction ID: Q1", b"aaab", b"abcd abbcd", b"", b"ORDER 12345 order 7"]
    return [dict(pattern=p, haystack=b64(h), spans=py_spans(p.encode(), h)) for p in pats for h in hays]


if __name__ == "__main__":
    with open(os.path.join(HERE, "emails_v1.json"), "w") as f:
        json.dump(emails(), f, indent=1)
    with open(os.path.join(HERE, "regex_v1.json"), "w") as f:
        json.dump(regexes(), f, indent=1)
    print("golden fixtures written")
