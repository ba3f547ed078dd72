// abi_pack.hpp — Solidity-ABI packer / reader for the verifier outputs (host code).
//
// Replaces, for whole batches, the step that follows the hot path for on-chain consumers:
//   core/src/io.rs:5-16   sol! { struct SolEmailOutput { bytes32 from_domain_hash; bytes32 public_key_hash;
//                                                       string[] external_inputs; }
//                                struct SolEmailWithRegexOutput { SolEmailOutput email; string[] matches; } }
//   core/src/io.rs:35-45  VerificationOutput::abi_encode  (SolValue::abi_encode of the struct = the value wrapped in a
//                         one-element sequence, i.e. what Solidity's abi.encode(structValue) returns)
//   helpers/src/io.rs:12-31  AbiDecodable::abi_decode: SolEmailOutput first, SolEmailWithRegexOutput second, both with
//                         validate = true (the bytes must be exactly the canonical encoding)
// alloy-sol-types is absent from /root/reference (Cargo.lock dependency); the layout below is the contract ABI
// specification's head/tail encoding, which is what that crate implements.
//
// Layout written here (every word 32 bytes, big-endian integers):
//   EmailOnly : [0x20] fdh pkh [0x60] S(external_inputs)
//   WithRegex : [0x20] [0x40] [0x40 + 0x60 + |S(ext)|] fdh pkh [0x60] S(ext) S(matches)
//   S(v)      : [n] [off_0 .. off_{n-1}] (len_i, bytes_i zero-padded to 32)...   off_i relative to the first offset word
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../include/zkemail_b200.h"

namespace zkb {
namespace abi {

inline size_t pad32(size_t n) { return (n + 31) & ~(size_t)31; }

inline size_t string_array_size(const zkb_str* v, size_t n) {
  size_t s = 32 + 32 * n;
  for (size_t i = 0; i < n; i++) s += 32 + pad32(v[i].len);
  return s;
}

inline size_t encoded_size(const zkb_output_view& o) {
  const size_t email = 96 + string_array_size(o.external_inputs, o.n_external_inputs);
  return o.with_regex ? 32 + 64 + email + string_array_size(o.matches, o.n_matches) : 32 + email;
}

inline void put_word(uint8_t* p, uint64_t v) {
  memset(p, 0, 24);
  for (int i = 0; i < 8; i++) p[24 + i] = (uint8_t)(v >> (8 * (7 - i)));
}

inline uint8_t* put_string_array(uint8_t* p, const zkb_str* v, size_t n) {
  put_word(p, n);
  uint8_t* heads = p + 32;
  uint8_t* tail = heads + 32 * n;
  for (size_t i = 0; i < n; i++) {
    put_word(heads + 32 * i, (uint64_t)(tail - heads));
    put_word(tail, v[i].len);
    const size_t padded = pad32(v[i].len);
    if (v[i].len) memcpy(tail + 32, v[i].s, v[i].len);
    if (padded > v[i].len) memset(tail + 32 + v[i].len, 0, padded - v[i].len);
    tail += 32 + padded;
  }
  return tail;
}

inline uint8_t* put_email(uint8_t* p, const zkb_output_view& o) {
  memcpy(p, o.from_domain_hash, 32);
  memcpy(p + 32, o.public_key_hash, 32);
  put_word(p + 64, 0x60);
  return put_string_array(p + 96, o.external_inputs, o.n_external_inputs);
}

// writes exactly encoded_size(o) bytes
inline void encode(const zkb_output_view& o, uint8_t* p) {
  put_word(p, 0x20);
  if (!o.with_regex) { put_email(p + 32, o); return; }
  const size_t email = 96 + string_array_size(o.external_inputs, o.n_external_inputs);
  put_word(p + 32, 0x40);
  put_word(p + 64, 0x40 + email);
  uint8_t* q = put_email(p + 96, o);
  put_string_array(q, o.matches, o.n_matches);
}

// ---- reader.  Permissive structural parse with bounds checks, then the canonical re-encoding must equal
// the input byte for byte — the same acceptance set as a validating decode.
struct Reader {
  const uint8_t* d;
  size_t n;
  bool word(size_t at, uint64_t& v) const {   // a word that must fit 63 bits (every offset / length here)
    if (at > n || n - at < 32) return false;
    for (int i = 0; i < 24; i++) if (d[at + i]) return false;
    v = 0;
    for (int i = 0; i < 8; i++) v = (v << 8) | d[at + 24 + i];
    return v < ((uint64_t)1 << 62);
  }
};

inline bool valid_utf8(const uint8_t* s, size_t n) {
  size_t i = 0;
  while (i < n) {
    const uint8_t c = s[i];
    if (c < 0x80) { i++; continue; }
    size_t need; uint32_t cp, lo;
    if ((c & 0xE0) == 0xC0) { need = 1; cp = c & 0x1F; lo = 0x80; }
    else if ((c & 0xF0) == 0xE0) { need = 2; cp = c & 0x0F; lo = 0x800; }
    else if ((c & 0xF8) == 0xF0) { need = 3; cp = c & 0x07; lo = 0x10000; }
    else return false;
    if (n - i <= need) return false;
    for (size_t k = 1; k <= need; k++) { if ((s[i + k] & 0xC0) != 0x80) return false; cp = (cp << 6) | (s[i + k] & 0x3F); }
    if (cp < lo || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return false;
    i += need + 1;
  }
  return true;
}

inline bool read_string_array(const Reader& r, size_t at, std::vector<zkb_span>& out) {
  uint64_t cnt;
  if (!r.word(at, cnt) || cnt > r.n / 32) return false;
  const size_t heads = at + 32;
  for (uint64_t i = 0; i < cnt; i++) {
    uint64_t off, len;
    if (!r.word(heads + 32 * (size_t)i, off) || off > r.n) return false;
    const size_t sp = heads + (size_t)off;
    if (!r.word(sp, len) || len > r.n || sp + 32 > r.n || r.n - (sp + 32) < len) return false;
    if (!valid_utf8(r.d + sp + 32, (size_t)len)) return false;
    out.push_back(zkb_span{(uint64_t)(sp + 32), len});
  }
  return true;
}

inline bool read_email(const Reader& r, size_t at, zkb_abi_decoded& dec, std::vector<zkb_span>& spans) {
  if (at > r.n || r.n - at < 96) return false;
  memcpy(dec.from_domain_hash, r.d + at, 32);
  memcpy(dec.public_key_hash, r.d + at + 32, 32);
  uint64_t off;
  if (!r.word(at + 64, off) || off > r.n) return false;
  return read_string_array(r, at + (size_t)off, spans);
}

inline bool reencodes(const uint8_t* data, size_t len, const zkb_abi_decoded& dec, const std::vector<zkb_span>& spans) {
  std::vector<zkb_str> strs(spans.size());
  for (size_t i = 0; i < spans.size(); i++) { strs[i].s = (const char*)data + spans[i].off; strs[i].len = (size_t)spans[i].len; }
  zkb_output_view v;
  v.from_domain_hash = dec.from_domain_hash; v.public_key_hash = dec.public_key_hash;
  v.external_inputs = strs.data(); v.n_external_inputs = dec.n_external_inputs;
  v.matches = strs.data() + dec.n_external_inputs; v.n_matches = dec.n_matches;
  v.with_regex = dec.with_regex;
  if (encoded_size(v) != len) return false;
  std::vector<uint8_t> again(len);
  encode(v, again.data());
  return memcmp(again.data(), data, len) == 0;
}

// helpers/src/io.rs:12-31 — EmailOnly is tried first, WithRegex second
inline bool decode(const uint8_t* data, size_t len, zkb_abi_decoded& dec, std::vector<zkb_span>& spans) {
  Reader r{data, len};
  uint64_t top;
  if (!r.word(0, top) || top > len) return false;
  const size_t T = (size_t)top;
  {
    memset(&dec, 0, sizeof dec);
    spans.clear();
    if (read_email(r, T, dec, spans)) {
      dec.with_regex = 0; dec.n_external_inputs = (uint32_t)spans.size(); dec.n_matches = 0;
      if (reencodes(data, len, dec, spans)) return true;
    }
  }
  memset(&dec, 0, sizeof dec);
  spans.clear();
  uint64_t oe, om;
  if (!r.word(T, oe) || !r.word(T + 32, om) || oe > len || om > len) return false;
  if (!read_email(r, T + (size_t)oe, dec, spans)) return false;
  dec.n_external_inputs = (uint32_t)spans.size();
  if (!read_string_array(r, T + (size_t)om, spans)) return false;
  dec.n_matches = (uint32_t)(spans.size() - dec.n_external_inputs);
  dec.with_regex = 1;
  return reencodes(data, len, dec, spans);
}

}  // namespace abi
}  // namespace zkb
