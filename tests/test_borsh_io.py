"""borsh encodings of the input structs (core/src/structs.rs:1-62 under the `risc0` feature).  The expected bytes are
written out by hand from the borsh specification (u32 LE lengths, Option tag byte, usize as u64 LE, fields in order)."""
import pytest

import zkemail_rs_b200 as z
from zkemail_rs_b200.structs import CompiledRegex, DFA, RegexInfo


def test_email_bytes_are_the_specified_layout():
    email = z.Email("a.b", b"\x01\x02\xff", z.PublicKey(b"\x30\x00", "rsa"),
                    [z.ExternalInput("n", "v1", 300), z.ExternalInput("m", None, 0)])
    want = (b"\x03\x00\x00\x00a.b" + b"\x03\x00\x00\x00\x01\x02\xff" + b"\x02\x00\x00\x00\x30\x00" + b"\x03\x00\x00\x00rsa"
            + b"\x02\x00\x00\x00"
            + b"\x01\x00\x00\x00n" + b"\x01" + b"\x02\x00\x00\x00v1" + (300).to_bytes(8, "little")
            + b"\x01\x00\x00\x00m" + b"\x00" + (0).to_bytes(8, "little"))
    assert z.to_borsh(email) == want
    assert z.from_borsh(z.Email, want) == email


def test_email_with_regex_round_trip_and_strictness():
    email = z.Email("ex.com", b"From: x\r\n\r\nhi", z.PublicKey(b"\x30\x03\x02\x01\x05", "rsa"), [])
    ewr = z.EmailWithRegex(email, RegexInfo([CompiledRegex(DFA(b"\x01\x02", b"\x03"), ["cap", "é"]), CompiledRegex(DFA(b"", b""), None)], None))
    b = z.to_borsh(ewr)
    assert z.from_borsh(z.EmailWithRegex, b) == ewr
    # header_parts: Some(vec of 2) ... body_parts: None is the last byte
    assert b[-1] == 0 and b.count(b"\x03\x00\x00\x00cap") == 1
    for bad in (b[:-1], b + b"\x00", b[:-1] + b"\x02"):
        with pytest.raises(z.BorshError):
            z.from_borsh(z.EmailWithRegex, bad)
    with pytest.raises(z.BorshError):
        z.from_borsh(z.Email, b"\x02\x00\x00\x00\xff\xfe" + b"\x00" * 12)      # invalid UTF-8 in from_domain
    with pytest.raises(z.BorshError):
        z.from_borsh(z.Email, b"\xff\xff\xff\xff")                              # absurd length


@pytest.mark.gpu
def test_borsh_inputs_verify_like_the_structs(engine):
    from tests.util import mixed_emails
    emails, _ = mixed_emails(seed=71, n_pos=12)
    blobs = [z.to_borsh(e) for e in emails]
    back = [z.from_borsh(z.Email, b) for b in blobs]
    assert engine.verify_batch(back).tobytes() == engine.verify_batch(emails).tobytes()
