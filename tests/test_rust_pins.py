"""Outputs of the REAL Rust reference for the committed golden vectors (tests/golden/rust_pins.json, written by
tools/pin_with_rust on a machine with a Rust toolchain).  When the file is present the oracle, the host front end,
the regex-automata wire reader, the regex compiler and the ABI packer are held to it; when it is absent (no toolchain
in the build image, no network) the tests skip with the reason — parity with the Rust binary then rests on the RFC
vectors, the real regex-automata blobs and the differential tests."""
import base64
import json
import os

import pytest

import oracle
import zkemail_rs_b200 as z
from tests.util import GOLDEN, NOW

PINS = os.path.join(GOLDEN, "rust_pins.json")
d64 = base64.b64decode
needs_pins = pytest.mark.skipif(not os.path.exists(PINS), reason="parity unpinned against the Rust binary: run tools/pin_with_rust "
                                                                 "(needs cargo + network) to create tests/golden/rust_pins.json")


def _vectors():
    out = {}
    for f in ("emails_v1.json", "rfc_vectors.json"):
        for g in json.load(open(os.path.join(GOLDEN, f))):
            out[g["name"]] = g
    return out


@needs_pins
def test_oracle_matches_the_rust_reference_on_every_golden_email():
    pins, vec = json.load(open(PINS)), _vectors()
    for p in pins["emails"]:
        g = vec[p["name"]]
        e = z.Email(g["from_domain"], d64(g["raw_email"]), z.PublicKey(d64(g["key"]), g["key_type"]))
        r = oracle.verify_email(e, NOW)
        if g["key_type"] == "ed25519":
            continue                                   # documented gap (ZKB_ST_UNSUPPORTED)
        assert (r["status"] == 0) == (p["status"] == "ok"), (p["name"], r["status"], p.get("panic_message"))
        if p["status"] == "ok":
            assert r["from_domain_hash"].hex() == p["from_domain_hash"] and r["public_key_hash"].hex() == p["public_key_hash"]
            out = z.VerificationOutput.from_parts(z.EmailVerifierOutput(r["from_domain_hash"], r["public_key_hash"], []), None)
            assert out.abi_encode().hex() == p["abi_encode"], p["name"]
        if "canon_header" in p:
            for impl in (oracle.canonicalize_signed_email, z.canonicalize_signed_email):
                hdr, body = impl(d64(g["raw_email"]), NOW)
                assert hdr == d64(p["canon_header"]) and body == d64(p["canon_body"]), (p["name"], impl.__module__)


@needs_pins
def test_wire_reader_and_compiler_match_regex_automata():
    from zkemail_rs_b200.engine import regex_automata_to_zdf
    pins = json.load(open(PINS))
    for p in pins["regex"]:
        if "error" in p:
            with pytest.raises(z.RegexError):
                z.compile_regex(p["pattern"])
            continue
        fwd, bwd = regex_automata_to_zdf(d64(p["fwd"]), False), regex_automata_to_zdf(d64(p["bwd"]), True)   # real blobs load
        ours = z.compile_regex(p["pattern"])
        for h in p["haystacks"]:
            hay, want = d64(h["haystack"]), [tuple(s) for s in h["spans"]]
            cnt, spans = oracle.dfa_find_iter(fwd, bwd, hay, 256)
            assert cnt == len(want) and list(spans) == want, ("real tables", p["pattern"])
            cnt, spans = oracle.dfa_find_iter(ours.fwd, ours.bwd, hay, 256)
            assert cnt == len(want) and list(spans) == want, ("compiled tables", p["pattern"])


@pytest.mark.gpu
@needs_pins
def test_engine_matches_the_rust_reference(engine):
    pins, vec = json.load(open(PINS)), _vectors()
    ps = [p for p in pins["emails"] if vec[p["name"]]["key_type"] != "ed25519"]
    emails = [z.Email(vec[p["name"]]["from_domain"], d64(vec[p["name"]]["raw_email"]),
                      z.PublicKey(d64(vec[p["name"]]["key"]), vec[p["name"]]["key_type"])) for p in ps]
    for p, r in zip(ps, engine.verify_batch(emails)):
        assert (int(r["status"]) == 0) == (p["status"] == "ok"), p["name"]
        if p["status"] == "ok":
            assert bytes(r["from_domain_hash"]).hex() == p["from_domain_hash"] and bytes(r["public_key_hash"]).hex() == p["public_key_hash"]
