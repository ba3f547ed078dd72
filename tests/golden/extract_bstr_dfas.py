"""Extracts REAL regex-automata dense-DFA serialisations (wire version 2, little-endian) from a binary that
statically links the `bstr` crate (>= 1.0), e.g. a ripgrep build:  bstr embeds
`whitespace_anchored_fwd.littleendian.dfa` and `whitespace_anchored_rev.littleendian.dfa`, which its build
script generates with `regex-cli generate serialize dense dfa` (pattern `\\s+`; anchored start kind; the
reverse one with match kind "all" — the same configuration dfa::regex::Builder uses for `DFA.bwd`,
helpers/src/regex.rs:7-14).  These are bytes written by the crate's own `write_to`, i.e. third-party pins of
the layout csrc/ra_wire.hpp reads.  (The reference's Cargo.lock pins regex-automata 0.4.9; the blobs carry
the same label and format version 2.)

usage: python tests/golden/extract_bstr_dfas.py /path/to/binary
writes tests/golden/ra_dense_ws_fwd.bin and tests/golden/ra_dense_ws_rev.bin (found in that order: bstr's
fsm modules are linked alphabetically)."""
import mmap
import os
import struct
import sys

LABEL = b"rust-regex-automata-dfa-dense\0\0\0"
HERE = os.path.dirname(os.path.abspath(__file__))


def blob_len(m, base):
    """Walks the sections of dense::DFA::write_to and returns the total length (or None)."""
    u32 = lambda o: struct.unpack_from("<I", m, o)[0]  # noqa: E731
    o = base + 32
    if u32(o) != 0xFEFF or u32(o + 4) != 2:
        return None
    o += 16                                   # endianness, version, unused, flags
    states, stride2 = u32(o), u32(o + 4)
    if not (0 < states < 1 << 20 and 1 <= stride2 <= 9):
        return None
    o += 8 + 256 + 4 * (states << stride2)    # byte classes, transitions
    o += 4 + 256                              # start kind, start byte map
    stride, plen = u32(o), u32(o + 4)
    o += 16 + 4 * (2 * stride + (0 if plen == 0xFFFFFFFF else stride * plen))
    ms = u32(o)
    o += 4 + 8 * ms
    o += 8 + 4 * u32(o + 4)                   # pattern_len, id_len, ids
    o += 32                                   # special
    o += 4 + 8 * u32(o)                       # accelerators
    return o + 32 - base                      # quit set


def main(path):
    with open(path, "rb") as f:
        m = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        found, at = [], 0
        while True:
            at = m.find(LABEL, at)
            if at < 0:
                break
            n = blob_len(m, at) if m[at + 32:at + 34] == b"\xff\xfe" else None
            if n:
                found.append(bytes(m[at:at + n]))
            at += 1
    assert len(found) == 2, f"expected bstr's two dense DFAs, found {len(found)}"
    for name, blob in zip(("ra_dense_ws_fwd.bin", "ra_dense_ws_rev.bin"), found):
        with open(os.path.join(HERE, name), "wb") as f:
            f.write(blob)
        print(name, len(blob), "bytes")


if __name__ == "__main__":
    main(sys.argv[1])
