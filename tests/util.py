"""Shared helpers for the test-suite: seeded key pool, synthetic mail sets, record comparison."""
import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import zkemail_rs_b200 as z  # noqa: E402
from zkemail_rs_b200 import synth  # noqa: E402

NOW = 1704067200
GOLDEN = os.path.join(ROOT, "tests", "golden")


@functools.lru_cache(maxsize=None)
def key_pool():
    """Keys are cached on disk as private DER so every run (CPU box, GPU box) signs identically."""
    out = {}
    for bits, count in ((2048, 3), (1024, 2)):
        ks = []
        for i in range(count):
            path = os.path.join(GOLDEN, f"key_{bits}_{i}.der")
            if os.path.exists(path):
                ks.append(synth.KeyPair.from_private_der(open(path, "rb").read()))
            else:
                k = synth.KeyPair.generate(bits)
                os.makedirs(GOLDEN, exist_ok=True)
                with open(path, "wb") as f:
                    f.write(k.private_der())
                ks.append(k)
        out[bits] = ks
    return out


def mixed_emails(seed=1, n_pos=24, with_token=False):
    """Positives over canonicalisation modes / sizes / key sizes, then every negative class."""
    rng = np.random.default_rng(seed)
    keys = key_pool()
    emails, labels = [], []
    sizes = [0, 1, 2, 55, 56, 63, 64, 65, 119, 512, 1000, 4096, 4097, 10000]
    canons = ["relaxed/relaxed", "simple/simple", "relaxed/simple", "simple/relaxed", "relaxed", "simple"]
    for i in range(n_pos):
        bits = 1024 if i % 5 == 4 else 2048
        k = keys[bits][i % len(keys[bits])]
        dom = f"mail{i % 3}.example.com"
        kw = dict(canon=canons[i % len(canons)], sig_position="top" if i % 2 == 0 else "bottom")
        if i % 7 == 3:
            kw["omit_c"] = True
            kw["canon"] = "simple/simple"
        bl = sizes[i % len(sizes)]
        if with_token:
            bl = max(bl, 400)
            kw["token"] = b"Transaction ID: A%07dZ" % i
            kw["qp_soft_breaks"] = (i % 3 == 0)
            kw["canon"] = "relaxed/relaxed"
            kw.pop("omit_c", None)
        emails.append(synth.make_email(rng, k, dom, idx=i, body_len=bl, **kw))
        labels.append("pos")
    base = [synth.make_email(rng, keys[2048][0], "mail0.example.com", idx=100 + j, body_len=700,
                             **({"token": b"Transaction ID: B%07dQ" % j} if with_token else {}))
            for j in range(len(synth.NEGATIVE_KINDS))]
    for j, kind in enumerate(synth.NEGATIVE_KINDS):
        emails.append(synth.mutate(base[j], kind, rng, other_key=keys[2048][1]))
        labels.append(kind)
    return emails, labels


def assert_records_equal(got, exp, label=""):
    """got: numpy record of the engine; exp: oracle dict.  Bit-exact on every field."""
    assert int(got["status"]) == exp["status"], (label, "status", int(got["status"]), exp["status"], int(got["dkim_detail"]), exp["dkim_detail"])
    assert int(got["dkim_detail"]) == exp["dkim_detail"], (label, "dkim_detail", int(got["dkim_detail"]), exp["dkim_detail"])
    for f in ("body_hash", "header_hash", "from_domain_hash", "public_key_hash"):
        assert bytes(got[f]) == exp[f], (label, f, bytes(got[f]).hex(), exp[f].hex())
    assert int(got["bh_ok"]) == exp["bh_ok"], (label, "bh_ok")
    assert int(got["rsa_ok"]) == exp["rsa_ok"], (label, "rsa_ok")
    parts = [tuple(int(x) for x in got["parts"][i]) for i in range(int(got["n_parts"]))]
    eparts = [tuple(p) for p in exp["parts"]]
    # start/end are only defined when there is at least one match
    norm = lambda ps: [(c, s, e, k) if c else (c, 0, 0, k) for (c, s, e, k) in ps]
    assert norm(parts) == norm(eparts), (label, "parts", parts, eparts)


def contiguous_views(emails):
    """Copies the raw messages of `emails` into ONE numpy buffer (so that it can be registered with the
    engine for the zero-copy / device-canonicalisation path) and returns (buffer, EmailViews)."""
    from zkemail_rs_b200.engine import EmailViews
    import ctypes as C
    offs, cur = [], 7
    for e in emails:
        offs.append(cur)
        cur += len(e.raw_email) + 13          # odd gaps: bodies start at arbitrary alignments
    buf = np.full(cur + 64, 0x20, dtype=np.uint8)
    keep = [buf]
    v = np.zeros((len(emails), 8), dtype=np.uint64)
    for i, e in enumerate(emails):
        raw = np.frombuffer(e.raw_email, dtype=np.uint8)
        buf[offs[i]:offs[i] + len(raw)] = raw
        dom = e.from_domain.encode()
        key = bytes(e.public_key.key)
        kt = e.public_key.key_type.encode()
        keep += [dom, key, kt]
        v[i] = [C.cast(C.c_char_p(dom), C.c_void_p).value or 0, len(dom), buf.ctypes.data + offs[i], len(raw),
                C.cast(C.c_char_p(key), C.c_void_p).value or 0, len(key), C.cast(C.c_char_p(kt), C.c_void_p).value or 0, len(kt)]
    return buf, EmailViews.from_arrays(v, keep)
