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
inline bool build_dfa_blob(const uint8_t* zdf, size_t len, bool want_reverse, std::vector<uint8_t>& blob,
                           uint32_t& elem_bytes) {
  ZdfView d;
  if (!zdf_parse(zdf, len, d)) return false;
  if (((d.flags & 1u) != 0) != want_reverse) return false;
  const uint64_t cells = (uint64_t)d.n_states * d.n_classes;
  if (cells >= (1ull << 31)) return false;
  elem_bytes = cells <= 65536 ? 2 : 4;
  size_t bytes = ZKB_DFA_HDR + cells * elem_bytes;
  blob.assign((bytes + 15) & ~(size_t)15, 0);
  uint32_t* hdr = reinterpret_cast<uint32_t*>(blob.data());
  hdr[0] = d.n_states; hdr[1] = d.n_classes;
  if (d.min_match <= d.max_match && d.max_match < d.n_states) {
    hdr[2] = d.min_match * d.n_classes; hdr[3] = d.max_match * d.n_classes;
  } else {  // no match states
    hdr[2] = 1; hdr[3] = 0;
  }
  hdr[4] = d.flags; hdr[5] = elem_bytes;
  for (int i = 0; i < 12; i++) hdr[6 + i] = d.start[i] * d.n_classes;
  memcpy(blob.data() + 128, d.class_map, 256);
  memcpy(blob.data() + 384, d.start_map, 256);
  if (elem_bytes == 2) {
    uint16_t* t = reinterpret_cast<uint16_t*>(blob.data() + ZKB_DFA_HDR);
    for (uint64_t i = 0; i < cells; i++) t[i] = (uint16_t)(rd32le(d.trans + 4 * i) * d.n_classes);
  } else {
    uint32_t* t = reinterpret_cast<uint32_t*>(blob.data() + ZKB_DFA_HDR);
    for (uint64_t i = 0; i < cells; i++) t[i] = rd32le(d.trans + 4 * i) * d.n_classes;
  }
  return true;
}

}  // namespace zkb
