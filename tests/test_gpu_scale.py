"""Parity at scale: tens of thousands of generator-made emails (mixed key sizes and body lengths, planted negatives,
regex token with soft breaks) through every input path of the engine — pageable multi-chunk pipeline, registered
memory, resident batch — against the multi-threaded oracle, field by field over the whole result arrays."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
import zkemail_rs_b200 as z
import workload as gen
from zkemail_rs_b200.engine import EmailViews, RESULT_DTYPE
from zkemail_rs_b200.structs import CompiledRegex, RegexInfo
from tests.util import NOW

pytestmark = pytest.mark.gpu

HEADER = [r"from:[^\r\n]*@d[0-9]+\.example\.com", r"\r\nsubject:[^\r\n]+\r\n"]
BODY = [r"Transaction ID: [A-Z0-9]+"]


def _oracle(pool, parts):
    views7 = pool.oracle_views()
    n = views7.shape[0]
    out = (oracle.Result * n)()
    keep = oracle._Keep()
    hp, nh = oracle._marshal_parts(parts[0], keep) if parts else (None, 0)
    bp, nb = oracle._marshal_parts(parts[1], keep) if parts else (None, 0)
    L = oracle.lib()
    rc = L.zo_verify_batch_mt(C.cast(views7.ctypes.data, C.POINTER(oracle._Email)), n, hp, nh, bp, nb, NOW,
                              os.cpu_count() or 1, 1 if L.zo_has_openssl() else 0, out)
    assert rc == 0
    return np.frombuffer(out, dtype=RESULT_DTYPE, count=n)


def _same(got, exp, what):
    for f in ("status", "dkim_detail", "body_hash", "header_hash", "from_domain_hash", "public_key_hash", "bh_ok", "rsa_ok", "n_parts"):
        bad = np.nonzero((got[f] != exp[f]).reshape(len(got), -1).any(axis=1))[0]
        assert bad.size == 0, (what, f, bad[:5], got[f][bad[:1]], exp[f][bad[:1]])
    for p in range(int(exp["n_parts"].max(initial=0))):
        live = exp["n_parts"] > p
        assert np.array_equal(got["parts"][live, p, :3], exp["parts"][live, p, :3]), (what, "parts", p)


@pytest.mark.parametrize("with_regex", [False, True])
def test_large_mixed_batch_all_input_paths(with_regex):
    n = 40000
    rng = np.random.default_rng(77)
    keys = gen.KeyPool(24, 24, 0)
    body = np.exp(rng.uniform(np.log(300), np.log(20000), size=n)).astype(np.uint32)
    pool = gen.MailPool(keys, n, body, seed=91, neg_fraction=0.02, token=with_regex, qp_percent=15 if with_regex else 0)
    parts = ([CompiledRegex(z.compile_regex(p), None) for p in HEADER], [CompiledRegex(z.compile_regex(p), None) for p in BODY]) if with_regex else None
    exp = _oracle(pool, parts)
    assert int((exp["status"] == 0).sum()) == int(pool.expected_ok().sum())
    eng = z.Engine(device=0, now_unix=NOW, chunk_emails=3000)      # ~14 chunks through the three-slot pipeline
    try:
        rs = z.RegexSet(eng, RegexInfo(*parts)) if with_regex else None
        views = EmailViews.from_arrays(pool.engine_views(), keep=pool)
        _same(eng.verify_views(views, rs, with_captures=False), exp, "pageable")
        assert eng.last_batch_bytes()["host_front_end_emails"] == 0
        eng.register_host(pool.raw)
        try:
            _same(eng.verify_views(views, rs, with_captures=False), exp, "registered")
        finally:
            eng.unregister_host(pool.raw)
        pb = eng.prepare(views, rs, with_captures=False)
        pb.run_async(); pb.run_async()
        _same(pb.fetch(), exp, "resident")
        pb.close()
    finally:
        eng.close()


def test_oversized_messages_take_the_same_paths():
    """Bodies larger than a pinned staging block (8 MiB), next to ordinary mail, pageable and registered."""
    from zkemail_rs_b200 import synth
    from tests.util import assert_records_equal, contiguous_views, key_pool
    rng = np.random.default_rng(5)
    k = key_pool()[2048][0]
    big = bytes(rng.integers(32, 127, 9 * 1024 * 1024 + 17, dtype=np.uint8)).replace(b"  ", b" \r\n")
    emails = [synth.make_email(rng, k, "big.example.com", idx=1, body_len=500),
              z.Email("big.example.com", synth.sign_email([("From", "a@big.example.com"), ("Subject", "big one")], big, k, "big.example.com"), z.PublicKey(k.der, "rsa")),
              synth.make_email(rng, k, "big.example.com", idx=2, body_len=70000, canon="simple/simple")]
    emails.append(z.Email("big.example.com", emails[1].raw_email[:-5] + b"XXXXX", emails[1].public_key))   # body hash mismatch
    exp = oracle.verify_batch(emails, now=NOW)
    assert [e["status"] for e in exp] == [0, 0, 0, 3]
    eng = z.Engine(device=0, now_unix=NOW)
    try:
        for i, g in enumerate(eng.verify_batch(emails)):
            assert_records_equal(g, exp[i], ("pageable", i))
        buf, views = contiguous_views(emails)
        eng.register_host(buf)
        try:
            for i, g in enumerate(eng.verify_views(views)):
                assert_records_equal(g, exp[i], ("registered", i))
        finally:
            eng.unregister_host(buf)
    finally:
        eng.close()
