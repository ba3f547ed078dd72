"""helpers/src/generator.rs row: offline generate_email_inputs /
generate_email_with_regex_inputs.  The CPU half checks the host pieces (key records, signature listing,
soft-break cleaner against the oracle); the GPU half checks candidate walking against a restatement of the
reference loop that uses the oracle as its verifier."""
import base64

import numpy as np
import pytest

import oracle
import zkemail_rs_b200 as z
from zkemail_rs_b200 import generator as G
from zkemail_rs_b200 import synth
from zkemail_rs_b200.structs import RegexConfig, RegexPattern
from tests.util import NOW, key_pool


def _txt(der: bytes, k="rsa") -> str:
    return f"v=DKIM1; k={k}; p=" + base64.b64encode(der).decode()


def test_key_record_parsing_matches_cryptography():
    from cryptography.hazmat.primitives import serialization
    from cryptography.hazmat.primitives.asymmetric import rsa
    from cryptography.hazmat.primitives.serialization import load_der_public_key
    for bits in (1024, 2048):
        k = key_pool()[bits][0]
        pub = load_der_public_key(_spki(k.der))
        spki = pub.public_bytes(serialization.Encoding.DER, serialization.PublicFormat.SubjectPublicKeyInfo)
        pkcs1 = pub.public_bytes(serialization.Encoding.DER, serialization.PublicFormat.PKCS1)
        assert pkcs1 == k.der
        assert z.parse_dkim_key_record(_txt(spki)) == (k.der, "rsa")          # SPKI, the usual DNS form
        assert z.parse_dkim_key_record(_txt(k.der)) == (k.der, "rsa")         # bare PKCS#1 is accepted too
        assert z.parse_dkim_key_record("p=" + base64.b64encode(spki).decode()) == (k.der, "rsa")   # k= defaults to rsa
    assert z.parse_dkim_key_record(_txt(b"\x07" * 32, "ed25519")) == (b"\x07" * 32, "ed25519")
    for bad in ("v=DKIM1; k=rsa", "v=DKIM1; k=rsa; p=", _txt(b"\x01" * 31, "ed25519"), _txt(b"junk"), "k=dsa; p=AAAA", "p=@@@"):
        with pytest.raises(z.GeneratorError):
            z.parse_dkim_key_record(bad)


def _spki(pkcs1: bytes) -> bytes:
    bit = b"\x03" + G._der_len(len(pkcs1) + 1) + b"\x00" + pkcs1
    body = G._RSA_ALGID + bit
    return b"\x30" + G._der_len(len(body)) + body


def test_signature_listing_and_soft_break_cleaner():
    rng = np.random.default_rng(2)
    k = key_pool()[2048][0]
    a = synth.make_email(rng, k, "a.example.com", idx=1, body_len=120).raw_email
    b = synth.make_email(rng, k, "B.Example.com", idx=2, body_len=120, selector="s2").raw_email
    sig_b = b[: b.find(b"Received:")]
    broken = a[: a.find(b"Received:")].replace(b" v=1;", b" v=2;", 1)
    assert z.dkim_signatures(sig_b + broken + a) == [("B.Example.com", "s2"), None, ("a.example.com", "sel1")]
    assert z.dkim_signatures(b"Subject: x\r\n\r\nbody") == []
    with pytest.raises(z.GeneratorError):
        z.dkim_signatures(b" leading space\r\n\r\nbody")
    for body in (b"", b"=\r\n", b"ab=\r\ncd=\r\n=\r\nx", b"=\r=\r\n\n=", bytes(rng.integers(0, 256, 4000, dtype=np.uint8)).replace(b"\x00", b"=\r\n")):
        cleaned, imap = z.remove_quoted_printable_soft_breaks(body)
        exp, kept = oracle.qp_clean(body)
        assert cleaned == exp and len(cleaned) == len(body)
        assert int((imap != np.iinfo(np.uint64).max).sum()) == kept
        assert all(body[int(imap[i])] == cleaned[i] for i in range(kept))


def _reference_generate(from_domain, raw, keys):
    """helpers/src/generator.rs:11-53 with the oracle as verify_email_with_key."""
    sigs = z.dkim_signatures(raw, NOW)
    if not sigs:
        return "No DKIM signatures found"
    for ds in sigs:
        if ds is None or ds[0].lower() != from_domain.lower():
            continue
        got = G._resolve(keys, from_domain, ds[1])
        if got is None:
            continue
        e = z.Email(from_domain, raw, z.PublicKey(*got))
        if oracle.verify_batch([e], now=NOW)[0]["status"] == 0:
            return e
    return "No valid DKIM key found for any signature"


@pytest.mark.gpu
def test_generate_email_inputs_walks_candidates_like_the_reference(engine):
    rng = np.random.default_rng(9)
    k0, k1, k2 = key_pool()[2048][0], key_pool()[2048][1], key_pool()[1024][0]
    a1 = synth.make_email(rng, k0, "a.example.com", idx=1, body_len=300, selector="s-old").raw_email
    a2 = synth.make_email(rng, k1, "a.example.com", idx=1, body_len=300, selector="s-new").raw_email
    other = synth.make_email(rng, k2, "b.example.com", idx=2, body_len=300, selector="sb").raw_email
    hdr = lambda raw: raw[: raw.find(b"Received:")]  # noqa: E731
    keys = z.StaticKeys({
        ("a.example.com", "s-old"): _txt(_spki(k0.der)),
        ("a.example.com", "s-new"): (k1.der, "rsa"),
        ("a.example.com", "s-wrong"): _txt(_spki(k2.der)),
        ("a.example.com", "s-ed"): _txt(b"\x05" * 32, "ed25519"),
        ("a.example.com", "s-empty"): "v=DKIM1; k=rsa; p=",
        ("b.example.com", "sb"): (k2.der, "rsa"),
    })
    wrong_sel = a1.replace(b"s=s-old;", b"s=s-wrong;", 1)   # signature no longer verifies, and its key is another one
    ed_sel = hdr(a1).replace(b"s=s-old;", b"s=s-ed;", 1)
    empty_sel = hdr(a1).replace(b"s=s-old;", b"s=s-empty;", 1)
    unknown_sel = hdr(a1).replace(b"s=s-old;", b"s=s-nokey;", 1)
    items = [
        ("a.example.com", a1), ("A.EXAMPLE.COM", a1), ("a.example.com", a2),
        ("a.example.com", hdr(other) + a1),                    # foreign-domain header first
        ("a.example.com", unknown_sel + ed_sel + empty_sel + a1),   # three unusable candidates, the fourth passes
        ("a.example.com", wrong_sel),                          # only candidate fails
        ("b.example.com", a1),                                 # no header for that domain
        ("a.example.com", b"Subject: none\r\n\r\nbody"),       # no signatures at all
        ("a.example.com", synth.mutate(z.Email("a.example.com", a1, z.PublicKey(k0.der, "rsa")), "body_flip", rng).raw_email),
        ("b.example.com", hdr(a1) + other),
    ]
    got = z.generate_email_inputs_batch(items, keys, engine=engine)
    for (dom, raw), g in zip(items, got):
        exp = _reference_generate(dom, raw, keys)
        if isinstance(exp, str):
            assert isinstance(g, z.GeneratorError) and str(g) == exp, (dom, exp, g)
        else:
            assert isinstance(g, z.Email) and g == exp
    assert [isinstance(g, z.Email) for g in got] == [True, True, True, True, True, False, False, False, False, True]
    with pytest.raises(z.GeneratorError, match="No DKIM signatures found"):
        z.generate_email_inputs("a.example.com", items[7][1], keys, engine=engine)
    ext = [z.ExternalInput("addr", "0x1", 42)]
    assert z.generate_email_inputs("a.example.com", a1, keys, ext, engine).external_inputs == ext


@pytest.mark.gpu
def test_generate_email_with_regex_inputs_round_trip(engine):
    rng = np.random.default_rng(10)
    k = key_pool()[2048][0]
    e = synth.make_email(rng, k, "shop.example.com", idx=4, body_len=900, token=b"Transaction ID: Q0012345Z", qp_soft_breaks=True)
    keys = z.StaticKeys({("shop.example.com", "sel1"): _txt(_spki(k.der))})
    cfg = RegexConfig(header_parts=[RegexPattern(r"\r\nsubject:([^\r\n]+)\r\n", [1])],
                      body_parts=[RegexPattern(r"Transaction ID: ([A-Z0-9]+)", [1])])
    inp = z.generate_email_with_regex_inputs("shop.example.com", e.raw_email, cfg, keys, engine=engine)
    assert inp.regex_info.body_parts[0].captures == ["Q0012345Z"]
    assert len(inp.regex_info.header_parts[0].captures) == 1
    out = engine.verify_email_with_regex(inp)                     # the generated inputs feed the hot path
    assert out.regex_matches == inp.regex_info.header_parts[0].captures + ["Q0012345Z"]
    blob = z.VerificationOutput.from_output(out).abi_encode()     # and the packer after it
    assert z.abi_decode(blob).matches == out.regex_matches
    with pytest.raises(z.RegexError):
        z.generate_email_with_regex_inputs("shop.example.com", e.raw_email,
                                           RegexConfig(body_parts=[RegexPattern(r"[A-Z]", None)]), keys, engine=engine)
    empty = z.generate_email_with_regex_inputs("shop.example.com", e.raw_email, RegexConfig(header_parts=[], body_parts=None), keys, engine=engine)
    assert empty.regex_info.header_parts is None and empty.regex_info.body_parts is None


@pytest.mark.gpu
def test_example_script_on_files(tmp_path):
    """examples/verify_eml.py: .eml + key records + RegexConfig JSON in, serde JSON + ABI bytes out."""
    import json
    import os
    import subprocess
    import sys
    rng = np.random.default_rng(12)
    k = key_pool()[2048][1]
    e = synth.make_email(rng, k, "files.example.com", idx=5, body_len=600, token=b"Transaction ID: F00D1234")
    (tmp_path / "m.eml").write_bytes(e.raw_email)
    (tmp_path / "keys.json").write_text(json.dumps({"sel1": _txt(_spki(k.der))}))
    (tmp_path / "regex.json").write_text(json.dumps({"header_parts": None, "body_parts": [{"pattern": "Transaction ID: ([A-Z0-9]+)", "capture_indices": [1]}]}))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "examples", "verify_eml.py"), str(tmp_path / "m.eml"), "files.example.com",
                          str(tmp_path / "keys.json"), str(tmp_path / "regex.json")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.strip().splitlines()
    res = z.from_serde(z.EmailWithRegexVerifierOutput, json.loads(lines[0]))
    assert res.regex_matches == ["F00D1234"]
    exp = oracle.verify_batch([e], now=0)[0]
    assert res.email.from_domain_hash == exp["from_domain_hash"] and res.email.public_key_hash == exp["public_key_hash"]
    blob = bytes.fromhex(lines[1].split("abi:", 1)[1].strip())
    assert z.abi_decode(blob).matches == ["F00D1234"]
