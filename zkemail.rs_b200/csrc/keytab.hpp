// keytab.hpp — host-side public-key handling: strict PKCS#1 RSAPublicKey DER decode with the
// rsa 0.9.6 acceptance rules (DkimPublicKey::try_from_bytes, core/src/email.rs:28-29;
// SURVEY.md A.2 "Key load rejects ..."), and the per-key Montgomery constants the RSA kernel
// needs (n, R^2 mod n, -n^-1 mod 2^32).  Computed once per unique key and cached.
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace zkb {

struct RsaKeyInfo {
  std::vector<uint32_t> n;  // little-endian limbs, minimal
  uint64_t e = 0;
  uint32_t bits = 0, k = 0;   // modulus bits, byte length
  uint32_t limbs_class = 0;   // 32 / 64 / 128
};

namespace der {
inline bool read_len(const uint8_t* p, size_t n, size_t& pos, size_t& out) {
  if (pos >= n) return false;
  uint8_t b = p[pos++];
  if (b < 0x80) { out = b; return true; }
  int k = b & 0x7f;
  if (k == 0 || k > 4 || pos + k > n) return false;
  size_t v = 0;
  for (int i = 0; i < k; i++) v = (v << 8) | p[pos++];
  if (v < 0x80) return false;
  if (k > 1 && (v >> (8 * (k - 1))) == 0) return false;
  out = v;
  return true;
}
inline bool read_uint(const uint8_t* p, size_t n, size_t& pos, const uint8_t*& v, size_t& vl) {
  if (pos >= n || p[pos++] != 0x02) return false;
  size_t l;
  if (!read_len(p, n, pos, l) || l == 0 || pos + l > n) return false;
  const uint8_t* q = p + pos;
  if (q[0] & 0x80) return false;
  if (l > 1 && q[0] == 0 && !(q[1] & 0x80)) return false;
  if (l > 1 && q[0] == 0) { q++; l--; }
  v = q; vl = l;
  pos = (size_t)(q - p) + l;
  return true;
}
}  // namespace der

// Returns false when rsa::RsaPublicKey::from_pkcs1_der + check_public would reject the key.
inline bool parse_rsa_public_key(const uint8_t* d, size_t len, RsaKeyInfo& out) {
  size_t pos = 0, sl;
  if (len < 2 || d[pos++] != 0x30) return false;
  if (!der::read_len(d, len, pos, sl) || pos + sl != len) return false;
  const uint8_t *nv, *ev;
  size_t nl, el;
  if (!der::read_uint(d, len, pos, nv, nl)) return false;
  if (!der::read_uint(d, len, pos, ev, el)) return false;
  if (pos != len) return false;
  if (nl > 512 + 1) return false;
  uint32_t bits = 0;
  if (!(nl == 1 && nv[0] == 0)) {
    bits = (uint32_t)(nl - 1) * 8;
    for (uint8_t t = nv[0]; t; t >>= 1) bits++;
  }
  if (bits > 4096 || bits == 0) return false;
  if (el > 8) return false;
  uint64_t e = 0;
  for (size_t i = 0; i < el; i++) e = (e << 8) | ev[i];
  if (!(nv[nl - 1] & 1)) return false;  // n even
  // e >= n ?
  if (bits <= 64) {
    uint64_t nn = 0;
    for (size_t i = 0; i < nl; i++) nn = (nn << 8) | nv[i];
    if (e >= nn) return false;
  }
  if (!(e & 1)) return false;
  if (e < 2) return false;
  if (e > ((1ull << 33) - 1)) return false;
  out.bits = bits;
  out.k = (bits + 7) / 8;
  out.e = e;
  uint32_t limbs = (bits + 31) / 32;
  out.n.assign(limbs, 0);
  for (size_t i = 0; i < nl; i++) {
    size_t bi = nl - 1 - i;
    out.n[bi / 4] |= (uint32_t)nv[i] << (8 * (bi % 4));
  }
  out.limbs_class = limbs <= 32 ? 32 : limbs <= 64 ? 64 : 128;  // kernel instantiations
  return true;
}

// Fills one key-table entry (ZKB_KEY_STRIDE words).
inline void build_key_entry(const RsaKeyInfo& key, uint32_t* ent) {
  memset(ent, 0, sizeof(uint32_t) * ZKB_KEY_STRIDE);
  const int L = (int)key.limbs_class;
  for (size_t i = 0; i < key.n.size(); i++) ent[i] = key.n[i];
  // n0inv = -n^-1 mod 2^32 (Newton)
  uint32_t n0 = key.n[0], inv = 1;
  for (int i = 0; i < 5; i++) inv *= 2u - n0 * inv;
  ent[ZKB_KEY_N0INV] = 0u - inv;
  // R^2 mod n by 2*32*L modular doublings of 1 (done once per unique key)
  std::vector<uint32_t> x(L + 1, 0), t(L + 1, 0);
  x[0] = 1;
  auto geq = [&](const std::vector<uint32_t>& a) {  // a >= n (a has L+1 limbs)
    if (a[L]) return true;
    for (int i = L - 1; i >= 0; i--) {
      uint32_t ni = ent[i];
      if (a[i] != ni) return a[i] > ni;
    }
    return true;
  };
  for (int it = 0; it < 2 * 32 * L; it++) {
    uint32_t c = 0;
    for (int i = 0; i <= L; i++) {
      uint32_t v = x[i];
      x[i] = (v << 1) | c;
      c = v >> 31;
    }
    if (geq(x)) {
      uint64_t b = 0;
      for (int i = 0; i <= L; i++) {
        uint64_t ni = i < L ? ent[i] : 0;
        uint64_t d = (uint64_t)x[i] - ni - b;
        x[i] = (uint32_t)d;
        b = (d >> 32) & 1;
      }
    }
  }
  for (int i = 0; i < L; i++) ent[ZKB_KEY_RR + i] = x[i];
  ent[ZKB_KEY_ELO] = (uint32_t)key.e;
  ent[ZKB_KEY_EHI] = (uint32_t)(key.e >> 32);
  ent[ZKB_KEY_K] = key.k;
  ent[ZKB_KEY_LIMBS] = key.limbs_class;
}

}  // namespace zkb
