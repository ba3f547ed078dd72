"""regex-automata wire-format reader.

Pins: tests/golden/ra_dense_ws_{fwd,rev}.bin are dense DFAs serialised by the regex-automata crate itself (the
`\\s+` tables bstr embeds; provenance in tests/golden/extract_bstr_dfas.py).  The reader must load them, a table walk
with the search semantics of SURVEY.md R5 (one-byte match delay, EOI class, anchored start by look-behind kind) must
reproduce Unicode White_Space on them, the test writer (oracle/ra_wire.py) must re-serialise them byte for byte, and
the real reverse table must work as `DFA.bwd` in the scan kernel.  The remaining tests write the repo's own tables
in that (now pinned) layout and require identical scans."""
import os
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
            wire = W.zdf_to_wire(blob)
            assert wire.startswith(b"rust-regex-automata-dfa-dense\0") and len(wire) % 4 == 0
            back = regex_automata_to_zdf(wire, rev)
            a, b = W.parse_zdf(back), W.parse_zdf(blob)
            if not (b["mn"] <= b["mx"] < b["ns"]):
                b["mn"], b["mx"] = 1, 0
            assert a == b, pat
            # with the unreachable quit state the crate always emits, states shift by one but the automaton is the same
            wq = regex_automata_to_zdf(W.zdf_to_wire(blob, with_quit_state=True), rev)
            assert W.parse_zdf(wq)["ns"] == W.parse_zdf(blob)["ns"] + 1


REAL_FWD = open(os.path.join(os.path.dirname(__file__), "golden", "ra_dense_ws_fwd.bin"), "rb").read()
REAL_REV = open(os.path.join(os.path.dirname(__file__), "golden", "ra_dense_ws_rev.bin"), "rb").read()
# Unicode White_Space (what `\s` means in regex-syntax's Unicode mode)
WHITE_SPACE = [0x9, 0xA, 0xB, 0xC, 0xD, 0x20, 0x85, 0xA0, 0x1680, *range(0x2000, 0x200B), 0x2028, 0x2029, 0x202F, 0x205F, 0x3000]
WS_TEXTS = ["  \t\r\n x", "x ", "", " \u3000a", "".join(map(chr, WHITE_SPACE)), "\u200b ", "a \u2003\u2028", "\x1c ", " \x85\xa0.",
            "tail   ", "\u1680\u1681", "no-space", " \r\n\t\u2029"]


def _anchored_walk(p, hay, reverse):
    """Anchored search over a parsed ZDF1 table (regex-automata's loop: start state by look-behind kind — Text at
    either end of the haystack —, one transition per byte, a match state entered at byte i reports offset i, the EOI
    class after the last byte, dead state 0 stops).  Returns the last reported offset or None."""
    s, last, nc = p["start"][6 + 2], None, p["nc"]
    order = range(len(hay) - 1, -1, -1) if reverse else range(len(hay))
    for i in order:
        s = p["trans"][s * nc + p["classes"][hay[i]]]
        if p["mn"] <= s <= p["mx"]:
            last = i + 1 if reverse else i
        if s == 0:
            return last
    s = p["trans"][s * nc + nc - 1]
    if p["mn"] <= s <= p["mx"]:
        last = 0 if reverse else len(hay)
    return last


def test_real_crate_blobs_load_and_mean_white_space():
    assert len(REAL_FWD) == 2964 and len(REAL_REV) == 3232
    pf = W.parse_zdf(regex_automata_to_zdf(REAL_FWD, False))
    pr = W.parse_zdf(regex_automata_to_zdf(REAL_REV, True))
    assert (pf["ns"], pf["nc"], pf["flags"] & 6) == (17, 27, 2) and (pr["ns"], pr["nc"], pr["flags"]) == (19, 27, 3)
    assert pf["start"][:6] == [0] * 6 and pr["start"][:6] == [0] * 6          # anchored-only tables
    for t in WS_TEXTS:
        hay = t.encode()
        lead = next((i for i, ch in enumerate(t) if ord(ch) not in WHITE_SPACE), len(t))
        trail = next((i for i, ch in enumerate(reversed(t)) if ord(ch) not in WHITE_SPACE), len(t))
        exp_f = len(t[:lead].encode()) or None
        exp_r = len(hay) - len(t[len(t) - trail:].encode()) if trail else None
        assert _anchored_walk(pf, hay, False) == exp_f, (t, "fwd")
        assert _anchored_walk(pr, hay, True) == exp_r, (t, "rev")
    # truncated UTF-8 stops the match where the last whole scalar ended
    assert _anchored_walk(pf, b" \xe2\x80", False) == 1


def test_test_writer_reproduces_the_real_blobs_byte_for_byte():
    """parse (product reader) -> write (test writer) is the identity on crate output, so every other test in this
    file feeds the reader bytes in the crate's real layout."""
    assert W.zdf_to_wire(regex_automata_to_zdf(REAL_FWD, False), has_quit_state=True, start_kind=2) == REAL_FWD
    assert W.zdf_to_wire(regex_automata_to_zdf(REAL_REV, True), has_quit_state=True) == REAL_REV


def test_flags_word_is_one_bitset():
    # a three-word flags section (one u32 per flag) is NOT the crate's layout and is rejected
    three = REAL_FWD[:44] + struct.pack("<3I", 0, 1, 0) + REAL_FWD[48:]
    with pytest.raises(z.RegexError):
        regex_automata_to_zdf(three, False)


def _ws_pairs():
    ours = z.compile_regex(r"\s+")
    return ours, DFA(ours.fwd, REAL_REV)


def test_real_reverse_table_as_bwd_in_emulation():
    """`DFA.bwd` of the reference is an anchored, match-kind-all reverse DFA: exactly what the real blob is.  Paired
    with this repo's forward table for the same pattern it must give the same spans as the repo's own reverse table."""
    from tests import emu
    ours, mixed = _ws_pairs()
    hays = [t.encode() for t in WS_TEXTS] + [b"a  b", b"\r\n\r\n", "x\u3000\u3000y z".encode()]
    a = np.asarray(emu.dfa_scan(ours.fwd, ours.bwd, hays))
    b = np.asarray(emu.dfa_scan(mixed.fwd, mixed.bwd, hays))
    assert np.array_equal(a, b)
    assert any(int(r[0]) >= 1 and int(r[1]) > 0 for r in a)
    real_rev_zdf = regex_automata_to_zdf(REAL_REV, True)        # the oracle reads ZDF1 tables only
    for h, row in zip(hays, b):
        cnt, spans = oracle.dfa_find_iter(ours.fwd, real_rev_zdf, h)
        assert int(row[0]) == cnt and (cnt == 0 or (int(row[1]), int(row[2])) == tuple(spans[0]))


@pytest.mark.gpu
def test_real_reverse_table_as_bwd_on_the_device(engine):
    ours, mixed = _ws_pairs()
    hays = [t.encode() for t in WS_TEXTS] + [b"a  b", b"\r\n\r\n", "x\u3000\u3000y z".encode()] * 50
    assert np.array_equal(engine.dfa_scan_batch(ours, hays), engine.dfa_scan_batch(mixed, hays))


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
        wf, wb = W.zdf_to_wire(dfa.fwd, with_quit_state=True), W.zdf_to_wire(dfa.bwd)
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
