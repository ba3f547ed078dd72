"""TEST ORACLE (not product code): writer of regex-automata's dense-DFA serialisation, used to exercise the
library's reader (csrc/ra_wire.hpp) without a Rust toolchain.  It lays a ZDF1 table out the way
dense::DFA::to_bytes_little_endian does as restated in SURVEY.md §8a R5 (regex-automata 0.4.9, a Cargo.lock
dependency absent from /root/reference): label, endianness check, version, flags, transition table with
premultiplied ids, start table, match states, special states, accelerators, quit set.  The layout is PINNED by
two blobs written by the crate itself (tests/golden/ra_dense_ws_{fwd,rev}.bin, see extract_bstr_dfas.py):
re-serialising their parsed form through this writer reproduces them byte for byte (tests/test_ra_wire.py)."""
from __future__ import annotations

import struct

LABEL = b"rust-regex-automata-dfa-dense\0"
MAX = 0xFFFFFFFF


def parse_zdf(z: bytes) -> dict:
    magic, flags, ns, nc, mn, mx = struct.unpack_from("<6I", z, 0)
    assert magic == 0x3146445A
    start = list(struct.unpack_from("<12I", z, 24))
    trans = list(struct.unpack_from(f"<{ns * nc}I", z, 584))
    return dict(flags=flags, ns=ns, nc=nc, mn=mn, mx=mx, start=start, classes=z[72:328], start_map=z[328:584], trans=trans)


def zdf_to_wire(z: bytes, with_quit_state: bool = False, has_quit_state: bool = False, start_kind=None) -> bytes:
    """with_quit_state inserts an unreachable quit state as state 1 (what the crate's determinizer always emits);
    has_quit_state says state 1 of `z` already is that state.  start_kind: 0 both, 1 unanchored, 2 anchored
    (default: anchored for reverse tables, both otherwise — what dfa::regex::Builder produces)."""
    d = parse_zdf(z)
    ns, nc = d["ns"], d["nc"]
    stride2 = max(1, (nc - 1).bit_length())
    stride = 1 << stride2
    remap = list(range(ns))
    if with_quit_state:  # an unreachable quit state as state 1, like every DFA the crate's determinizer emits
        remap = [0] + [s + 1 for s in range(1, ns)]
        ns += 1
    sid = lambda s: remap[s] << stride2  # noqa: E731
    out = bytearray(LABEL + b"\0" * (-len(LABEL) % 4))
    out += struct.pack("<3I", 0xFEFF, 2, 0)
    has_empty, is_utf8 = (d["flags"] >> 2) & 1, (d["flags"] >> 1) & 1
    out += struct.pack("<I", has_empty | (is_utf8 << 1))
    out += struct.pack("<2I", ns, stride2) + d["classes"]
    table = [0] * (ns << stride2)
    for s in range(d["ns"]):
        for c in range(nc):
            table[sid(s) + c] = sid(d["trans"][s * nc + c])
    out += struct.pack(f"<{len(table)}I", *table)
    if start_kind is None:
        start_kind = 2 if d["flags"] & 1 else 0
    # a universal start state exists when the look-behind byte does not matter (all six kinds share one state)
    uni = [sid(h[0]) if len(set(h)) == 1 and h[0] != 0 else MAX for h in (d["start"][:6], d["start"][6:])]
    out += struct.pack("<I", start_kind) + d["start_map"] + struct.pack("<4I", 6, MAX, uni[0], uni[1])
    out += struct.pack("<12I", *[sid(s) for s in d["start"]])
    has_match = d["mn"] <= d["mx"] < d["ns"]
    n_match = d["mx"] - d["mn"] + 1 if has_match else 0
    out += struct.pack("<I", n_match)
    for i in range(n_match):
        out += struct.pack("<2I", i, 1)
    out += struct.pack("<2I", 1, n_match) + struct.pack(f"<{n_match}I", *([0] * n_match))
    mn, mx = (sid(d["mn"]), sid(d["mx"])) if has_match else (0, 0)
    quit_id = stride if (with_quit_state or has_quit_state) else 0
    out += struct.pack("<8I", max(mx, quit_id), quit_id, mn, mx, 0, 0, 0, 0)
    out += struct.pack("<I", 0) + b"\0" * 32
    return bytes(out)
