"""helpers/src/file.rs + the serde JSON shape of the API structs (core/src/structs.rs, helpers/src/structs.rs)."""
import json

import pytest

import zkemail_rs_b200 as z
from zkemail_rs_b200.structs import CompiledRegex, DFA, RegexConfig, RegexInfo, RegexPattern


def test_serde_shapes_round_trip(tmp_path):
    email = z.Email("example.com", b"From: a\r\n\r\nhi\xff", z.PublicKey(b"\x30\x03\x02\x01\x05", "rsa"),
                    [z.ExternalInput("addr", "0xabc", 42), z.ExternalInput("opt", None, 7)])
    v = z.to_serde(email)
    assert v == {"from_domain": "example.com", "raw_email": list(b"From: a\r\n\r\nhi\xff"),
                 "public_key": {"key": [0x30, 3, 2, 1, 5], "key_type": "rsa"},
                 "external_inputs": [{"name": "addr", "value": "0xabc", "max_length": 42}, {"name": "opt", "value": None, "max_length": 7}]}
    assert z.from_serde(z.Email, json.loads(json.dumps(v))) == email
    ewr = z.EmailWithRegex(email, RegexInfo([CompiledRegex(DFA(b"\x01\x02", b"\x03"), ["cap"])], None))
    p = tmp_path / "in.json"
    z.write_json_file(p, ewr)
    assert z.read_json_file(p, z.EmailWithRegex) == ewr
    out = z.EmailWithRegexVerifierOutput(z.EmailVerifierOutput(b"\x11" * 32, b"\x22" * 32, ["n", "v"]), ["m"])
    assert z.from_serde(z.EmailWithRegexVerifierOutput, z.to_serde(out)) == out


def test_regex_config_as_the_helpers_read_it(tmp_path):
    p = tmp_path / "regex.json"
    p.write_text('{"header_parts": [{"pattern": "subject:[^\\\\r\\\\n]+", "capture_indices": [1]}, {"pattern": "x"}], "body_parts": null, "extra": 1}')
    cfg = z.read_json_file(p, RegexConfig)
    assert cfg == RegexConfig([RegexPattern("subject:[^\\r\\n]+", [1]), RegexPattern("x", None)], None)


def test_serde_strictness_and_file_errors(tmp_path):
    with pytest.raises(ValueError, match="missing field `key_type`"):
        z.from_serde(z.PublicKey, {"key": [1]})
    with pytest.raises(ValueError, match="array of bytes"):
        z.from_serde(z.PublicKey, {"key": [256], "key_type": "rsa"})
    with pytest.raises(ValueError, match="array of bytes"):
        z.from_serde(z.PublicKey, {"key": "AAAA", "key_type": "rsa"})
    with pytest.raises(ValueError, match="unsigned integer"):
        z.from_serde(z.ExternalInput, {"name": "a", "value": None, "max_length": -1})
    with pytest.raises(ValueError, match="null where a value is required"):
        z.from_serde(z.Email, {"from_domain": None, "raw_email": [], "public_key": {"key": [], "key_type": ""}, "external_inputs": []})
    with pytest.raises(OSError, match="Failed to open email file"):
        z.read_email_file(tmp_path / "missing.eml")
    bad = tmp_path / "bad.json"
    bad.write_text("{not json")
    with pytest.raises(ValueError, match="Failed to parse JSON"):
        z.read_json_file(bad, RegexConfig)
    eml = tmp_path / "a.eml"
    eml.write_bytes(b"From: x\r\n\r\nbody")
    assert z.read_email_file(eml) == b"From: x\r\n\r\nbody"
