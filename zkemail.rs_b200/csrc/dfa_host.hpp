// dfa_host.hpp — host side of the DFA kernel: validation of ZDF1 tables (the wire form of
// `DFA.fwd` / `DFA.bwd`, include/zkemail_b200.h) and their conversion to the device blob that
// dfa.cuh scans from shared memory.  Validation failure is the engine's equivalent of
// dense::DFA::from_bytes(..).unwrap() panicking (core/src/regex.rs:32-33).
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/zkemail_b200.h"
#include "dfa.cuh"
#include "ra_wire.hpp"

namespace zkb {

struct ZdfView {
  uint32_t flags = 0, n_states = 0, n_classes = 0, min_match = 0, max_match = 0;
  uint32_t start[12] = {0};
  const uint8_t* class_map = nullptr;
  const uint8_t* start_map = nullptr;
  const uint8_t* trans = nullptr;
};

inline uint32_t rd32le(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

inline bool zdf_parse(const uint8_t* b, size_t n, ZdfView& d) {
  if (!b || n < ZKB_ZDF_HEADER || rd32le(b) != ZKB_ZDF_MAGIC) return false;
  d.flags = rd32le(b + 4); d.n_states = rd32le(b + 8); d.n_classes = rd32le(b + 12);
  d.min_match = rd32le(b + 16); d.max_match = rd32le(b + 20);
  for (int i = 0; i < 12; i++) d.start[i] = rd32le(b + 24 + 4 * i);
  d.class_map = b + 72; d.start_map = b + 328; d.trans = b + ZKB_ZDF_HEADER;
  if (d.n_states == 0 || d.n_classes < 2 || d.n_classes > 257) return false;
  if ((uint64_t)d.n_states * d.n_classes * 4 + ZKB_ZDF_HEADER != n) return false;
  for (int i = 0; i < 12; i++) if (d.start[i] >= d.n_states) return false;
  for (int i = 0; i < 256; i++) {
    if (d.class_map[i] >= d.n_classes - 1) return false;
    if (d.start_map[i] > 5) return false;
  }
  const uint64_t cells = (uint64_t)d.n_states * d.n_classes;
  for (uint64_t i = 0; i < cells; i++) if (rd32le(d.trans + 4 * i) >= d.n_states) return false;
  return true;
}

// ZDF1 -> device blob (see dfa.cuh).  elem = 2 when every premultiplied id fits in 16 bits.
// direct = true expands the byte classes into 256-entry rows (+ an EOI column).
inline bool build_dfa_blob(const uint8_t* zdf, size_t len, bool want_reverse, bool direct, uint32_t force_elem,
                           std::vector<uint8_t>& blob, uint32_t& elem_bytes) {
  ZdfView d;
  if (!zdf_parse(zdf, len, d)) return false;
  if (((d.flags & 1u) != 0) != want_reverse) return false;
  const uint32_t stride = direct ? 256u : d.n_classes;
  const uint64_t cells = (uint64_t)d.n_states * stride;
  if (cells >= (1ull << 31)) return false;
  elem_bytes = force_elem ? force_elem : (cells <= 65536 ? 2 : 4);
  if (elem_bytes == 2 && cells > 65536) return false;
  size_t bytes = ZKB_DFA_HDR + (cells + (direct ? d.n_states : 0)) * elem_bytes;
  blob.assign((bytes + 15) & ~(size_t)15, 0);
  uint32_t* hdr = reinterpret_cast<uint32_t*>(blob.data());
  hdr[0] = d.n_states; hdr[1] = stride;
  // Device numbering: the dead state stays 0, match states go LAST, so the per-byte match test of the scan is one
  // compare (sid >= min_match).  perm[old] = new.
  // (the dead state is never a match: the search stops on it before the match test, whatever the range says)
  const uint32_t mlo = d.min_match ? d.min_match : 1u;
  const bool has_match = mlo <= d.max_match && d.max_match < d.n_states;
  std::vector<uint32_t> perm(d.n_states);
  {
    const uint32_t lo = has_match ? mlo : d.n_states, hi = has_match ? d.max_match : d.n_states;
    const uint32_t n_match = has_match ? hi - lo + 1 : 0;
    uint32_t next = 0;
    for (uint32_t s = 0; s < d.n_states; s++) {
      if (s >= lo && s <= hi) perm[s] = d.n_states - n_match + (s - lo);
      else perm[s] = next++;
    }
    if (has_match) { hdr[2] = (d.n_states - n_match) * stride; hdr[3] = (d.n_states - 1) * stride; }
    else { hdr[2] = 0xffffffffu; hdr[3] = 0xffffffffu; }   // no state id reaches this: never a match
  }
  hdr[4] = d.flags; hdr[5] = elem_bytes;
  for (int i = 0; i < 12; i++) hdr[6 + i] = perm[d.start[i]] * stride;
  hdr[18] = direct ? 1u : 0u;
  memcpy(blob.data() + 128, d.class_map, 256);
  memcpy(blob.data() + 384, d.start_map, 256);
  auto put = [&](uint64_t idx, uint32_t v) {
    if (elem_bytes == 2) reinterpret_cast<uint16_t*>(blob.data() + ZKB_DFA_HDR)[idx] = (uint16_t)v;
    else reinterpret_cast<uint32_t*>(blob.data() + ZKB_DFA_HDR)[idx] = v;
  };
  // state 0 is the dead state: the search stops there and its transitions are never consulted by the
  // reference algorithm; the kernel relies on it being absorbing, so its row is forced to zeros
  auto tr = [&](uint32_t s, uint32_t c) { return s == 0 ? 0u : perm[rd32le(d.trans + 4 * ((uint64_t)s * d.n_classes + c))] * stride; };
  if (!direct) {
    for (uint32_t s = 0; s < d.n_states; s++)
      for (uint32_t c = 0; c < d.n_classes; c++) put((uint64_t)perm[s] * stride + c, tr(s, c));
  } else {
    for (uint32_t s = 0; s < d.n_states; s++) {
      for (uint32_t b = 0; b < 256; b++) put((uint64_t)perm[s] * 256 + b, tr(s, d.class_map[b]));
      put(cells + perm[s], tr(s, d.n_classes - 1));
    }
  }
  return true;
}

// Chooses the table form for a forward/reverse pair: DIRECT rows while both tables stay small enough
// for >= 8 CTAs of shared memory per SM, one element width for both.
inline bool build_dfa_pair(const uint8_t* fwd, size_t fwd_len, const uint8_t* bwd, size_t bwd_len,
                           std::vector<uint8_t>& fb, std::vector<uint8_t>& rb, uint32_t& elem, bool& direct) {
  // regex-automata wire blobs (what a Rust caller's DFA.fwd / DFA.bwd hold) are converted to ZDF1 first
  std::vector<uint8_t> fz, rz;
  if (ra::is_wire(fwd, fwd_len)) { if (!ra::to_zdf(fwd, fwd_len, false, fz)) return false; fwd = fz.data(); fwd_len = fz.size(); }
  if (ra::is_wire(bwd, bwd_len)) { if (!ra::to_zdf(bwd, bwd_len, true, rz)) return false; bwd = rz.data(); bwd_len = rz.size(); }
  ZdfView f, r;
  if (!zdf_parse(fwd, fwd_len, f) || !zdf_parse(bwd, bwd_len, r)) return false;
  direct = ((uint64_t)f.n_states + r.n_states) * 257 * 2 <= 24 * 1024;
  const uint64_t fc = (uint64_t)f.n_states * (direct ? 256 : f.n_classes), rc = (uint64_t)r.n_states * (direct ? 256 : r.n_classes);
  elem = (fc <= 65536 && rc <= 65536) ? 2 : 4;
  uint32_t e1, e2;
  return build_dfa_blob(fwd, fwd_len, false, direct, elem, fb, e1) && build_dfa_blob(bwd, bwd_len, true, direct, elem, rb, e2);
}

}  // namespace zkb
