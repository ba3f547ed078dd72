"""core/src/io.rs / helpers/src/io.rs row: the library's batch ABI packer against the
oracle's generic ABI encoder, which is itself pinned by the worked examples of the contract ABI specification."""
import os

import pytest
from hypothesis import given, settings, strategies as st

import zkemail_rs_b200 as z
from oracle import abi_ref as A

W = lambda n: int(n).to_bytes(32, "big")  # noqa: E731
R = lambda b: b + b"\0" * (-len(b) % 32)  # noqa: E731


def test_oracle_matches_abi_specification_examples():
    # sam(bytes,bool,uint256[]) with ("dave", true, [1,2,3])
    got = A.enc_sequence([("bytes", b"dave"), ("bool", True), ("array", [("uint", 1), ("uint", 2), ("uint", 3)])])
    exp = W(0x60) + W(1) + W(0xA0) + W(4) + R(b"dave") + W(3) + W(1) + W(2) + W(3)
    assert got == exp
    # f(uint256,uint32[],bytes10,bytes) with (0x123, [0x456, 0x789], "1234567890", "Hello, world!")
    got = A.enc_sequence([("uint", 0x123), ("array", [("uint", 0x456), ("uint", 0x789)]), ("bytesN", b"1234567890"), ("bytes", b"Hello, world!")])
    exp = W(0x123) + W(0x80) + R(b"1234567890") + W(0xE0) + W(2) + W(0x456) + W(0x789) + W(13) + R(b"Hello, world!")
    assert got == exp
    # g(uint256[][],string[]) with ([[1, 2], [3]], ["one", "two", "three"])
    got = A.enc_sequence([("array", [("array", [("uint", 1), ("uint", 2)]), ("array", [("uint", 3)])]),
                          ("array", [("string", "one"), ("string", "two"), ("string", "three")])])
    exp = (W(0x40) + W(0x140) + W(2) + W(0x40) + W(0xA0) + W(2) + W(1) + W(2) + W(1) + W(3)
           + W(3) + W(0x60) + W(0xA0) + W(0xE0) + W(3) + R(b"one") + W(3) + R(b"two") + W(5) + R(b"three"))
    assert got == exp


def _out(fdh, pkh, ext, matches):
    return z.VerificationOutput.from_parts(z.EmailVerifierOutput(fdh, pkh, list(ext)), None if matches is None else list(matches))


_TXT = st.text(max_size=70)


@settings(max_examples=200, deadline=None)
@given(st.binary(min_size=32, max_size=32), st.binary(min_size=32, max_size=32), st.lists(_TXT, max_size=6),
       st.one_of(st.none(), st.lists(_TXT, max_size=5)))
def test_encode_matches_oracle_and_round_trips(fdh, pkh, ext, matches):
    o = _out(fdh, pkh, ext, matches)
    blob = o.abi_encode()
    assert blob == A.verification_output_abi_encode(fdh, pkh, ext, matches)
    back = z.VerificationOutput.abi_decode(blob)
    assert back == o


def test_known_layout_email_only():
    fdh, pkh = bytes(range(32)), bytes(range(32, 64))
    blob = _out(fdh, pkh, ["name", "value"], None).abi_encode()
    exp = W(0x20) + fdh + pkh + W(0x60) + W(2) + W(0x40) + W(0x80) + W(4) + R(b"name") + W(5) + R(b"value")
    assert blob == exp
    blob = _out(fdh, pkh, [], []).abi_encode()
    exp = W(0x20) + W(0x40) + W(0xC0) + fdh + pkh + W(0x60) + W(0) + W(0)
    assert blob == exp


def test_batch_encode_is_per_item_encode():
    outs = [_out(os.urandom(32), os.urandom(32), ["k%d" % i, "v" * (i % 70)], None if i % 3 else ["m" * (i % 40), "é%d" % i])
            for i in range(10000)]
    blobs = z.abi_encode_batch(outs)
    assert len(blobs) == len(outs)
    for i in (0, 1, 2, 31, 32, 33, 4095, 4096, 9999):
        o = outs[i]
        assert blobs[i] == A.verification_output_abi_encode(o.email.from_domain_hash, o.email.public_key_hash, o.email.external_inputs, o.matches)
    assert all(z.abi_decode(b) == o for b, o in zip(blobs[:500], outs[:500]))


def test_decode_rejects_what_a_validating_decode_rejects():
    fdh, pkh = bytes(range(32)), bytes(range(32, 64))
    good = _out(fdh, pkh, ["a", "bc"], ["m"]).abi_encode()
    only = _out(fdh, pkh, ["a", "bc"], None).abi_encode()
    bad = [
        b"", good[:-1], good[:-32], good + b"\0" * 32,                       # truncated / trailing data
        good[:31] + b"\x40" + good[32:],                                       # wrong top offset
        only[:32 + 64 + 31] + b"\x80" + only[32 + 64 + 32:],                   # non-canonical array offset
        only[:-1] + b"\x01",                                                   # dirty padding
        W(0x20) + fdh + pkh + W(0x60) + W(1) + W(0x20) + W(2) + R(b"\xff\xfe"),  # invalid UTF-8
        W(0x20) + fdh + pkh + W(0x60) + W(1 << 40),                            # absurd count
        W(0x20) + fdh + pkh + W(0x60) + W(1) + W(0x20) + W(1 << 50) + R(b"x"),  # absurd length
    ]
    for b in bad:
        with pytest.raises(z.AbiDecodeError):
            z.abi_decode(b)
    assert z.abi_decode(only).matches is None and z.abi_decode(good).matches == ["m"]


def test_engine_output_to_abi():
    """from_output accepts both reference output structs."""
    e = z.EmailVerifierOutput(b"\x11" * 32, b"\x22" * 32, ["n", "v"])
    assert z.VerificationOutput.from_output(e).matches is None
    r = z.EmailWithRegexVerifierOutput(e, ["x"])
    assert z.VerificationOutput.from_output(r).abi_encode() == A.verification_output_abi_encode(b"\x11" * 32, b"\x22" * 32, ["n", "v"], ["x"])
