// ra_wire.hpp — reader for regex-automata's dense-DFA serialisation.
//
// The reference ships `DFA.fwd` / `DFA.bwd` as `dense::DFA::to_bytes_little_endian()` output with the leading
// alignment padding removed (helpers/src/regex.rs:7-14) and loads them with `dense::DFA::from_bytes`
// (core/src/regex.rs:32-33).  regex-automata 0.4.9 is a Cargo.lock dependency absent from /root/reference and no
// Rust toolchain exists here; the layout below is restated from the crate's format and PINNED by two
// blobs the crate itself wrote (tests/golden/ra_dense_ws_{fwd,rev}.bin: the `\s+` dense DFAs the bstr crate embeds,
// format version 2; tests/test_ra_wire.py), which settled the flags word as ONE u32 bitset.  The reader validates
// every section, requires the sections to consume the blob exactly, and rejects anything else (ZKB_E_REGEX)
// instead of guessing.  It converts to the engine's own ZDF1 table (include/zkemail_b200.h), which models the same automaton: premultiplied ids become
// state indices, match states stay one contiguous range entered one byte late, the last alphabet class is EOI.
//
//   label   "rust-regex-automata-dfa-dense" NUL, zero-padded to a multiple of 4      (32 bytes)
//   u32     0xFEFF endianness check, u32 version (2), u32 unused
//   flags   u32 bitset (bit0 has_empty, bit1 is_utf8, bit2 always_start_anchored)
//   transitions  u32 state_len, u32 stride2, u8 classes[256], u32 next[state_len << stride2]
//   starts  u32 kind, u8 start_map[256], u32 stride (6), u32 pattern_len | MAX, u32 universal unanchored | MAX,
//           u32 universal anchored | MAX, u32 ids[2 * stride (+ stride * pattern_len)]
//   matches u32 state_len, (u32 off, u32 len)[state_len], u32 pattern_len, u32 id_len, u32 ids[id_len]
//   special u32 max, quit, min_match, max_match, min_accel, max_accel, min_start, max_start
//   accels  u32 count, 8 bytes each;  quit set: 32 bytes
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/zkemail_b200.h"

namespace zkb {
namespace ra {

static const char kLabel[] = "rust-regex-automata-dfa-dense";

inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

inline bool is_wire(const uint8_t* b, size_t n) {
  return b && n >= 32 && memcmp(b, kLabel, sizeof(kLabel)) == 0;   // includes the NUL
}

struct Cursor {
  const uint8_t* b; size_t n, at;
  bool u32(uint32_t& v) { if (n - at < 4) return false; v = rd32(b + at); at += 4; return true; }
  const uint8_t* bytes(size_t k) { if (n - at < k) return nullptr; const uint8_t* p = b + at; at += k; return p; }
};

inline void wr32(std::vector<uint8_t>& o, size_t at, uint32_t v) { for (int i = 0; i < 4; i++) o[at + i] = (uint8_t)(v >> (8 * i)); }

inline bool to_zdf(const uint8_t* b, size_t n, bool reverse, std::vector<uint8_t>& zdf) {
  if (!is_wire(b, n)) return false;
  Cursor c{b, n, 32};
  uint32_t endian, version, unused;
  if (!c.u32(endian) || !c.u32(version) || !c.u32(unused) || endian != 0xFEFFu || version != 2u) return false;
  uint32_t bits;
  if (!c.u32(bits) || bits > 7u) return false;
  const bool has_empty = bits & 1u, is_utf8 = bits & 2u;
  // transition table
  uint32_t state_len, stride2;
  if (!c.u32(state_len) || !c.u32(stride2) || state_len == 0 || stride2 < 1 || stride2 > 9 || state_len > (1u << 24)) return false;
  const uint8_t* classes = c.bytes(256);
  if (!classes) return false;
  uint32_t max_class = 0;
  for (int i = 0; i < 256; i++) max_class = classes[i] > max_class ? classes[i] : max_class;
  const uint32_t alphabet = max_class + 2;   // + EOI
  const uint32_t stride = 1u << stride2;
  if (alphabet > stride || alphabet > 257) return false;
  for (int i = 1; i < 256; i++) if (classes[i] < classes[i - 1] || classes[i] > classes[i - 1] + 1) return false;   // classes are contiguous byte ranges
  const uint64_t cells = (uint64_t)state_len << stride2;
  if (cells > (1ull << 30)) return false;
  const uint8_t* table = c.bytes((size_t)cells * 4);
  if (!table) return false;
  const uint32_t id_limit = (uint32_t)cells;
  auto valid_id = [&](uint32_t id) { return id < id_limit && (id & (stride - 1)) == 0; };
  // start table
  uint32_t kind, sstride, pattern_len, uni_un, uni_an;
  if (!c.u32(kind) || kind > 2u) return false;
  const uint8_t* start_map = c.bytes(256);
  if (!start_map) return false;
  for (int i = 0; i < 256; i++) if (start_map[i] > 5) return false;
  if (!c.u32(sstride) || !c.u32(pattern_len) || !c.u32(uni_un) || !c.u32(uni_an) || sstride != 6u) return false;
  if (pattern_len != 0xFFFFFFFFu && pattern_len > 1u) return false;   // the helpers build single-pattern regexes
  if ((uni_un != 0xFFFFFFFFu && !valid_id(uni_un)) || (uni_an != 0xFFFFFFFFu && !valid_id(uni_an))) return false;
  const uint32_t n_start = 12u + (pattern_len == 0xFFFFFFFFu ? 0u : 6u * pattern_len);
  const uint8_t* starts = c.bytes((size_t)n_start * 4);
  if (!starts) return false;
  for (uint32_t i = 0; i < n_start; i++) if (!valid_id(rd32(starts + 4 * i))) return false;
  // match states
  uint32_t m_states, m_patterns, m_ids;
  if (!c.u32(m_states) || m_states > state_len) return false;
  const uint8_t* slices = c.bytes((size_t)m_states * 8);
  if (!slices || !c.u32(m_patterns) || !c.u32(m_ids) || m_ids > (1u << 24)) return false;
  const uint8_t* pids = c.bytes((size_t)m_ids * 4);
  if (!pids) return false;
  for (uint32_t i = 0; i < m_states; i++) {
    const uint64_t off = rd32(slices + 8 * i), len = rd32(slices + 8 * i + 4);
    if (off + len > m_ids || len == 0) return false;
  }
  for (uint32_t i = 0; i < m_ids; i++) if (rd32(pids + 4 * i) >= (m_patterns ? m_patterns : 1u)) return false;
  // special states
  uint32_t sp[8];
  for (auto& v : sp) if (!c.u32(v)) return false;
  const uint32_t sp_max = sp[0], quit = sp[1], min_match = sp[2], max_match = sp[3];
  for (int i = 0; i < 8; i++) if (!valid_id(sp[i])) return false;
  if (min_match > max_match || max_match > sp_max) return false;
  if ((min_match == 0) != (max_match == 0)) return false;
  const uint32_t n_match = min_match == 0 ? 0u : ((max_match - min_match) >> stride2) + 1;
  if (n_match != m_states) return false;
  if (quit != 0 && quit != stride) return false;   // the quit state, when present, is state 1
  // accelerators, quit set
  uint32_t n_accel;
  if (!c.u32(n_accel) || n_accel > state_len || !c.bytes((size_t)n_accel * 8)) return false;
  const uint8_t* quitset = c.bytes(32);
  if (!quitset) return false;
  if (c.at != n) return false;   // every section accounted for, nothing left over
  // transitions: all valid ids; the quit state must be unreachable (a search that reaches it errors in the
  // reference: only heuristic Unicode word boundaries produce it, helpers/src/regex.rs never enables them)
  for (int i = 0; i < 32; i++) if (quitset[i]) return false;
  for (uint32_t s = 0; s < state_len; s++)
    for (uint32_t k = 0; k < alphabet; k++) {
      const uint32_t id = rd32(table + 4 * (((uint64_t)s << stride2) + k));
      if (!valid_id(id)) return false;
      if (quit != 0 && id == quit && s != (quit >> stride2)) return false;
    }
  // emit ZDF1
  zdf.assign(ZKB_ZDF_HEADER + (size_t)state_len * alphabet * 4, 0);
  wr32(zdf, 0, ZKB_ZDF_MAGIC);
  wr32(zdf, 4, (reverse ? 1u : 0u) | (is_utf8 ? 2u : 0u) | (has_empty ? 4u : 0u));
  wr32(zdf, 8, state_len);
  wr32(zdf, 12, alphabet);
  if (n_match) { wr32(zdf, 16, min_match >> stride2); wr32(zdf, 20, max_match >> stride2); }
  else { wr32(zdf, 16, 1); wr32(zdf, 20, 0); }
  for (int i = 0; i < 12; i++) wr32(zdf, 24 + 4 * i, rd32(starts + 4 * i) >> stride2);
  memcpy(zdf.data() + 72, classes, 256);
  memcpy(zdf.data() + 328, start_map, 256);
  for (uint32_t s = 0; s < state_len; s++)
    for (uint32_t k = 0; k < alphabet; k++)
      wr32(zdf, ZKB_ZDF_HEADER + 4 * ((size_t)s * alphabet + k), rd32(table + 4 * (((uint64_t)s << stride2) + k)) >> stride2);
  return true;
}

}  // namespace ra
}  // namespace zkb
