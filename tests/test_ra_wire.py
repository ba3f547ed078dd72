"""regex-automata wire-format reader: blobs laid out as dense::DFA::to_bytes_little_endian
(restated layout, oracle/ra_wire.py) load into the engine and scan exactly like the ZDF1 tables they came from."""
import struct

import numpy as np
import pytest

import oracle
import zkemail_rs_b200 as z
from oracle import ra_wire as W
from zkemail_rs_b200.engine import regex_automata_to_zdf
from zkemail_rs_b200.structs import DFA

PATTERNS = [r"Transaction ID: [A-Z0-9]+", r"\r\nsubject:[^\r\n]+\r\n", r"a*", r"(?i)from:[^\r\n]*@example\.com", r"é+|x", r"^$", r"[^a]"]
HAYS = [b"", b"aaab", b"Transaction ID: A12Z and Transaction ID: Q", b"to:x\r\nsubject: hi there\r\nfrom:Bob <b@EXAMPLE.com>\r\n",
        "ééx é".encode(), b"\xff\xfeabc", bytes(range(256))]


def test_reader_inverts_the_restated_layout():
    for pat in PATTERNS:
        dfa = z.compile_regex(pat)
        for blob, rev in ((dfa.fwd, False), (dfa.bwd, True)):
            for fw in (1, 3):
                wire = W.zdf_to_wire(blob, flag_words=fw)
                assert wire.startswith(b"rust-regex-automata-dfa-dense\0") and len(wire) % 4 == 0
                back = regex_automata_to_zdf(wire, rev)
                a, b = W.parse_zdf(back), W.parse_zdf(blob)
                if not (b["mn"] <= b["mx"] < b["ns"]):
                    b["mn"], b["mx"] = 1, 0
                assert a == b, pat
            # with the unreachable quit state the crate always emits, states shift by one but the automaton is the same
            wq = regex_automata_to_zdf(W.zdf_to_wire(blob, with_quit_state=True), rev)
            assert W.parse_zdf(wq)["ns"] == W.parse_zdf(blob)["ns"] + 1


def test_reader_rejects_corrupt_blobs():
    dfa = z.compile_regex(r"ab+c")
    wire = W.zdf_to_wire(dfa.fwd)
    bad = [wire[:-4], wire + b"\0\0\0\0", wire[:32] + struct.pack("<I", 0xFFFE0000) + wire[36:],      # length, endianness
           wire[:36] + struct.pack("<I", 3) + wire[40:],                                                 # version
           wire[:52] + struct.pack("<I", 11) + wire[56:],                                                # stride2
           b"rust-regex-automata-dfa-sparse\0\0" + wire[32:], wire[:40]]
    tpos = 56 + 256
    bad.append(wire[:tpos + 8] + struct.pack("<I", 0x7FFFFFF0) + wire[tpos + 12:])                     # transition out of range
    bad.append(wire[:tpos + 8] + struct.pack("<I", 1) + wire[tpos + 12:])                              # id not a multiple of the stride
    for b in bad:
        with pytest.raises(z.RegexError):
            regex_automata_to_zdf(b, False)
    assert regex_automata_to_zdf(wire, False)


def test_wire_blobs_scan_like_their_zdf_tables_in_emulation():
    from tests import emu
    for pat in PATTERNS:
        dfa = z.compile_regex(pat)
        wf, wb = W.zdf_to_wire(dfa.fwd, with_quit_state=True), W.zdf_to_wire(dfa.bwd, flag_words=3)
        for qp in (False, True):
            a = emu.dfa_scan(dfa.fwd, dfa.bwd, HAYS, qp=qp)
            b = emu.dfa_scan(wf, wb, HAYS, qp=qp)
            assert np.array_equal(np.asarray(a), np.asarray(b)), (pat, qp)
        for h, row in zip(HAYS, np.asarray(emu.dfa_scan(wf, wb, HAYS))):
            cnt, spans = oracle.dfa_find_iter(dfa.fwd, dfa.bwd, h)
            assert int(row[0]) == cnt
            if cnt == 1:
                assert (int(row[1]), int(row[2])) == tuple(spans[0])


@pytest.mark.gpu
def test_wire_blobs_through_the_engine(engine):
    from zkemail_rs_b200.structs import CompiledRegex, RegexInfo
    from tests.util import NOW, assert_records_equal, mixed_emails
    for pat in PATTERNS:
        dfa = z.compile_regex(pat)
        wire = DFA(W.zdf_to_wire(dfa.fwd, with_quit_state=True), W.zdf_to_wire(dfa.bwd, with_quit_state=True))
        assert np.array_equal(engine.dfa_scan_batch(dfa, HAYS), engine.dfa_scan_batch(wire, HAYS)), pat
    emails, _ = mixed_emails(seed=3, n_pos=12, with_token=True)
    zd = z.compile_regex(r"Transaction ID: [A-Z0-9]+")
    info_w = RegexInfo(None, [CompiledRegex(DFA(W.zdf_to_wire(zd.fwd), W.zdf_to_wire(zd.bwd)), None)])
    got = engine.verify_with_regex_batch(emails, info_w)
    exp = oracle.verify_batch(emails, None, [CompiledRegex(zd, None)], now=NOW)
    for i, (g, e) in enumerate(zip(got, exp)):
        assert_records_equal(g, e, i)
