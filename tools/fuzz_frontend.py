"""Differential fuzzer: device front end (pageable-staged and registered-memory paths, with their host fallback)
against the CPU oracle on randomly dirty, randomly mutated signed mail.  Needs a GPU.

    python tools/fuzz_frontend.py [n_emails] [seed]

Every record field must be identical.  Dirt: odd header names / folds / tabs / non-ASCII / bare CR / LF, all four
canonicalisation pairs, optional tags (i= q= t= x= l=), duplicate and foreign signatures, post-signing byte edits."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import zkemail_rs_b200 as z  # noqa: E402
from zkemail_rs_b200 import synth  # noqa: E402
from tests.util import NOW, assert_records_equal, contiguous_views, key_pool  # noqa: E402

NAMES = [b"From", b"from", b"FROM", b"To", b"Subject", b"Date", b"X-Test", b"Cc", b"Message-ID", b"Reply-To"]
VALS = [b"a", b"B", b" ", b"\t", b"  ", b"\r\n ", b"\r\n\t", b"x@y.z", b";", b"=", b"\xc3\xa9", b"\xff", b":", b"\"q\"", b"<a@b.example.com>"]
BODY = [b"line", b" ", b"\t", b"\r\n", b"\n", b"\r", b"=\r\n", b"  ", b"text text", b"", b"\r\n\r\n", b" \r\n", b"\xe2\x82\xac", b"." * 70]
CANON = ["relaxed/relaxed", "simple/simple", "relaxed/simple", "simple/relaxed", "relaxed", "simple"]
EXTRA = ["", "", "", " t=1700000000;", " i=@d.example.com;", " i=user@sub.d.example.com;", " q=dns/txt;", " l=10;", " l=0;", " x=9999999999;",
         " x=1;", " z=From:a|To:b;", " q=other;", " i=@elsewhere.org;"]


def _insert_after_headers(raw, extra):
    i = raw.find(b"\r\n\r\n")
    return raw if i < 0 else raw[:i + 2] + extra + raw[i + 2:]


def build(n, seed):
    rng = np.random.default_rng(seed)
    keys = key_pool()[1024][:4]
    pick = lambda xs: xs[int(rng.integers(0, len(xs)))]  # noqa: E731
    emails = []
    foreign = []   # signature headers of other domains (the reference skips them; the device may too)
    for j in range(6):
        o = synth.make_email(rng, pick(keys), f"other{j}.example.org", idx=j, body_len=40,
                             canon=pick(CANON[:4]), **({"algo": "rsa-sha1"} if j == 5 else {})).raw_email
        foreign.append(o[: o.find(b"\r\n", o.find(b"\tb=")) + 2])
    while len(emails) < n:
        k = pick(keys)
        dom = "d.example.com"
        hs = []
        for _ in range(int(rng.integers(1, 7))):
            v = b"".join(pick(VALS) for _ in range(int(rng.integers(0, 8)))).rstrip(b"\r\n\t ")
            hs.append((pick(NAMES).decode("latin1"), v.decode("latin1")))
        if not any(nm.lower() == "from" for nm, _ in hs):
            hs.append(("From", "x@d.example.com"))
        body = b"".join(pick(BODY) for _ in range(int(rng.integers(0, 16))))
        hsel = tuple(pick([("from", "to", "subject", "date"), ("from",), ("From", "from", "subject"), ("to", "from", "x-test", "cc"), ("subject", "from", "from")]))
        try:
            raw = synth.sign_email(hs, body, k, dom, canon=pick(CANON), h=hsel, extra_tags=pick(EXTRA),
                                   sig_position=pick(["top", "bottom"]), omit_c=bool(rng.integers(0, 6) == 0))
        except Exception:
            continue
        e = z.Email(dom, raw, z.PublicKey(k.der, "rsa"))
        r = int(rng.integers(0, 12))
        if r == 0 and emails:      # second (foreign or stale) signature header in front
            other = emails[int(rng.integers(0, len(emails)))].raw_email
            cut = other.find(b"\r\n", other.find(b"\tb=")) + 2
            if other.startswith(b"DKIM-Signature") and cut > 2:
                e = z.Email(dom, other[:cut] + raw, e.public_key)
        elif r == 1:               # random byte edit after signing
            b = bytearray(raw)
            i = int(rng.integers(0, len(b)))
            b[i] = int(rng.integers(0, 256))
            e = z.Email(dom, bytes(b), e.public_key)
        elif r == 2:               # delete a byte
            i = int(rng.integers(0, len(raw)))
            e = z.Email(dom, raw[:i] + raw[i + 1:], e.public_key)
        elif r == 3:
            e = z.Email(pick(["D.EXAMPLE.COM", "example.com", "x.d.example.com"]), raw, e.public_key)
        elif r == 4:
            e = z.Email(dom, raw, z.PublicKey(pick(keys).der, "rsa"))
        elif r in (5, 6):          # signature header(s) of other domains in front of / behind the real one
            f = b"".join(pick(foreign) for _ in range(int(rng.integers(1, 3))))
            e = z.Email(dom, f + raw if r == 5 else raw + b"", e.public_key) if r == 5 else z.Email(dom, _insert_after_headers(raw, f), e.public_key)
        emails.append(e)
    return emails


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    t0 = time.time()
    emails = build(n, seed)
    exp = oracle.verify_batch(emails, now=NOW, threads=os.cpu_count() or 1)
    eng = z.Engine(device=0, now_unix=NOW, chunk_emails=1024)
    got_p = eng.verify_batch(emails)
    fb_p = eng.last_batch_bytes()["host_front_end_emails"]
    buf, views = contiguous_views(emails)
    eng.register_host(buf)
    got_r = eng.verify_views(views)
    fb_r = eng.last_batch_bytes()["host_front_end_emails"]
    eng.unregister_host(buf)
    bad = 0
    for i, e in enumerate(exp):
        for name, got in (("pageable", got_p), ("registered", got_r)):
            try:
                assert_records_equal(got[i], e, (name, i))
            except AssertionError as err:
                bad += 1
                if bad <= 5:
                    print("MISMATCH", err, emails[i].raw_email[:600], file=sys.stderr)
    st = np.bincount([e["status"] for e in exp], minlength=10)
    print(f"fuzz: {n} emails seed {seed}: mismatches {bad}; oracle status histogram {st.tolist()}; "
          f"host-front-end fallbacks pageable {fb_p} registered {fb_r}; {time.time() - t0:.1f}s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
