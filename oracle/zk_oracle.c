/*
 * zk_oracle.c — CPU ORACLE (test infrastructure; see zk_oracle.h header comment).
 * "parity unpinned" w.r.t. the Rust binary: pinned by KATs + differential tests + tests/golden.
 *
 * Plain C restatement of the reference hot path:
 *   core/src/circuits.rs:9-68, core/src/email.rs:25-36,61-86, core/src/regex.rs:15-53,
 *   core/src/crypto.rs:3-7, and the behaviour of the un-vendored crates it calls
 *   ([dep-memory] items are isolated in named functions below).
 *
 * Deliberately naive (operation order of the reference kept, e.g. the repeated " \r\n" scan);
 * the product code in zkemail.rs_b200/csrc is an independent implementation.
 */
#include "zk_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef ZO_WITH_OPENSSL
#include <openssl/bn.h>
#include <openssl/evp.h>
#include <openssl/sha.h>
#endif

static __thread int g_use_openssl = 0;

/* ------------------------------------------------------------------ byte buffer */
typedef struct {
  uint8_t *p;
  size_t n, cap;
} buf_t;
static void buf_reserve(buf_t *b, size_t add) {
  if (b->n + add <= b->cap) return;
  size_t nc = b->cap ? b->cap * 2 : 256;
  while (nc < b->n + add) nc *= 2;
  b->p = (uint8_t *)realloc(b->p, nc);
  b->cap = nc;
}
static void buf_put(buf_t *b, const void *d, size_t n) {
  buf_reserve(b, n + 1);
  if (n) memcpy(b->p + b->n, d, n);
  b->n += n;
}
static void buf_putc(buf_t *b, uint8_t c) { buf_put(b, &c, 1); }
static void buf_free(buf_t *b) {
  free(b->p);
  b->p = NULL;
  b->n = b->cap = 0;
}

/* ------------------------------------------------------------------ SHA-256 (FIPS 180-4)
 * sha2 0.10.9 one-shot, core/src/crypto.rs:3-7 */
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
#define ROR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))
static void sha256_block(uint32_t st[8], const uint8_t *p) {
  uint32_t w[64];
  for (int i = 0; i < 16; i++)
    w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) |
           ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
  for (int i = 16; i < 64; i++) {
    uint32_t s0 = ROR(w[i - 15], 7) ^ ROR(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = ROR(w[i - 2], 17) ^ ROR(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
  for (int i = 0; i < 64; i++) {
    uint32_t S1 = ROR(e, 6) ^ ROR(e, 11) ^ ROR(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = h + S1 + ch + K256[i] + w[i];
    uint32_t S0 = ROR(a, 2) ^ ROR(a, 13) ^ ROR(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}
void zo_sha256(const uint8_t *data, size_t len, uint8_t out[32]) {
#ifdef ZO_WITH_OPENSSL
  if (g_use_openssl) {
    SHA256(data, len, out);
    return;
  }
#endif
  uint32_t st[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                    0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  size_t full = len / 64;
  for (size_t i = 0; i < full; i++) sha256_block(st, data + 64 * i);
  uint8_t tail[128];
  size_t rem = len - 64 * full;
  memset(tail, 0, sizeof tail);
  if (rem) memcpy(tail, data + 64 * full, rem);
  tail[rem] = 0x80;
  size_t tl = (rem + 9 <= 64) ? 64 : 128;
  uint64_t bits = (uint64_t)len * 8;
  for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
  sha256_block(st, tail);
  if (tl == 128) sha256_block(st, tail + 64);
  for (int i = 0; i < 8; i++) {
    out[4 * i] = (uint8_t)(st[i] >> 24);
    out[4 * i + 1] = (uint8_t)(st[i] >> 16);
    out[4 * i + 2] = (uint8_t)(st[i] >> 8);
    out[4 * i + 3] = (uint8_t)st[i];
  }
}

/* ------------------------------------------------------------------ base64 (base64 0.21.7
 * general_purpose::STANDARD: padded alphabet, canonical padding REQUIRED on decode, trailing
 * bits must be zero, no whitespace tolerated) */
static const char B64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
size_t zo_base64_encode(const uint8_t *in, size_t n, char *out) {
  size_t o = 0, i = 0;
  for (; i + 3 <= n; i += 3) {
    uint32_t v = ((uint32_t)in[i] << 16) | ((uint32_t)in[i + 1] << 8) | in[i + 2];
    out[o++] = B64[v >> 18]; out[o++] = B64[(v >> 12) & 63];
    out[o++] = B64[(v >> 6) & 63]; out[o++] = B64[v & 63];
  }
  if (n - i == 1) {
    uint32_t v = (uint32_t)in[i] << 16;
    out[o++] = B64[v >> 18]; out[o++] = B64[(v >> 12) & 63]; out[o++] = '='; out[o++] = '=';
  } else if (n - i == 2) {
    uint32_t v = ((uint32_t)in[i] << 16) | ((uint32_t)in[i + 1] << 8);
    out[o++] = B64[v >> 18]; out[o++] = B64[(v >> 12) & 63]; out[o++] = B64[(v >> 6) & 63];
    out[o++] = '=';
  }
  return o;
}
static int b64val(uint8_t c) {
  if (c >= 'A' && c <= 'Z') return c - 'A';
  if (c >= 'a' && c <= 'z') return c - 'a' + 26;
  if (c >= '0' && c <= '9') return c - '0' + 52;
  if (c == '+') return 62;
  if (c == '/') return 63;
  return -1;
}
long zo_base64_decode(const uint8_t *in, size_t n, uint8_t *out) {
  if (n % 4 != 0) return -1; /* canonical padding => length multiple of 4 */
  size_t o = 0;
  for (size_t i = 0; i < n; i += 4) {
    int last = (i + 4 == n);
    int a = b64val(in[i]), b = b64val(in[i + 1]);
    if (a < 0 || b < 0) return -1;
    if (last && in[i + 2] == '=' && in[i + 3] == '=') {
      if (b & 15) return -1; /* non-zero trailing bits */
      out[o++] = (uint8_t)((a << 2) | (b >> 4));
    } else if (last && in[i + 3] == '=') {
      int c = b64val(in[i + 2]);
      if (c < 0 || (c & 3)) return -1;
      out[o++] = (uint8_t)((a << 2) | (b >> 4));
      out[o++] = (uint8_t)((b << 4) | (c >> 2));
    } else {
      int c = b64val(in[i + 2]), d = b64val(in[i + 3]);
      if (c < 0 || d < 0) return -1;
      out[o++] = (uint8_t)((a << 2) | (b >> 4));
      out[o++] = (uint8_t)((b << 4) | (c >> 2));
      out[o++] = (uint8_t)((c << 6) | d);
    }
  }
  return (long)o;
}

/* ------------------------------------------------------------------ bignum (32-bit limbs, LE)
 * num-bigint-dig 0.8.4 modpow restated as plain square-and-multiply with schoolbook multiply and
 * Knuth Algorithm D division — intentionally NOT Montgomery, so the GPU path is checked by a
 * different algorithm. */
#define BN_MAX 130 /* limbs: 4096-bit + slack */
typedef struct {
  uint32_t d[2 * BN_MAX + 2];
  int n; /* used limbs, no leading zero limbs (n==0 => zero) */
} bn_t;
static void bn_trim(bn_t *a) {
  while (a->n > 0 && a->d[a->n - 1] == 0) a->n--;
}
static void bn_from_be(bn_t *a, const uint8_t *p, size_t len) {
  memset(a, 0, sizeof *a);
  for (size_t i = 0; i < len; i++) {
    size_t bi = len - 1 - i; /* byte significance */
    if (bi / 4 < 2 * BN_MAX) a->d[bi / 4] |= (uint32_t)p[i] << (8 * (bi % 4));
  }
  a->n = (int)((len + 3) / 4);
  if (a->n > 2 * BN_MAX) a->n = 2 * BN_MAX;
  bn_trim(a);
}
static int bn_cmp(const bn_t *a, const bn_t *b) {
  if (a->n != b->n) return a->n < b->n ? -1 : 1;
  for (int i = a->n - 1; i >= 0; i--)
    if (a->d[i] != b->d[i]) return a->d[i] < b->d[i] ? -1 : 1;
  return 0;
}
static int bn_bits(const bn_t *a) {
  if (a->n == 0) return 0;
  uint32_t t = a->d[a->n - 1];
  int b = 0;
  while (t) { b++; t >>= 1; }
  return (a->n - 1) * 32 + b;
}
static void bn_mul(bn_t *r, const bn_t *a, const bn_t *b) {
  bn_t t;
  memset(&t, 0, sizeof t);
  for (int i = 0; i < a->n; i++) {
    uint64_t c = 0;
    for (int j = 0; j < b->n; j++) {
      uint64_t v = (uint64_t)a->d[i] * b->d[j] + t.d[i + j] + c;
      t.d[i + j] = (uint32_t)v;
      c = v >> 32;
    }
    t.d[i + b->n] = (uint32_t)c;
  }
  t.n = a->n + b->n;
  bn_trim(&t);
  *r = t;
}
/* r = a mod m (Knuth D), m != 0 */
static void bn_mod(bn_t *r, const bn_t *a, const bn_t *m) {
  if (bn_cmp(a, m) < 0) { *r = *a; return; }
  int n = m->n, mm = a->n - m->n;
  if (n == 1) {
    uint64_t rem = 0;
    for (int i = a->n - 1; i >= 0; i--) rem = ((rem << 32) | a->d[i]) % m->d[0];
    memset(r, 0, sizeof *r);
    r->d[0] = (uint32_t)rem; r->n = 1; bn_trim(r);
    return;
  }
  int s = 0;
  { uint32_t t = m->d[n - 1]; while (!(t & 0x80000000u)) { t <<= 1; s++; } }
  static __thread uint32_t u[2 * BN_MAX + 3], v[BN_MAX + 1];
  for (int i = n - 1; i > 0; i--) v[i] = s ? (m->d[i] << s) | (m->d[i - 1] >> (32 - s)) : m->d[i];
  v[0] = m->d[0] << s;
  u[a->n] = s ? a->d[a->n - 1] >> (32 - s) : 0;
  for (int i = a->n - 1; i > 0; i--) u[i] = s ? (a->d[i] << s) | (a->d[i - 1] >> (32 - s)) : a->d[i];
  u[0] = a->d[0] << s;
  for (int j = mm; j >= 0; j--) {
    uint64_t num = ((uint64_t)u[j + n] << 32) | u[j + n - 1];
    uint64_t qhat = num / v[n - 1], rhat = num % v[n - 1];
    while (qhat >= (1ull << 32) || qhat * v[n - 2] > ((rhat << 32) | u[j + n - 2])) {
      qhat--; rhat += v[n - 1];
      if (rhat >= (1ull << 32)) break;
    }
    int64_t borrow = 0; uint64_t carry = 0;
    for (int i = 0; i < n; i++) {
      uint64_t p = qhat * v[i] + carry;
      carry = p >> 32;
      int64_t t = (int64_t)u[i + j] - borrow - (int64_t)(p & 0xffffffffu);
      u[i + j] = (uint32_t)t;
      borrow = (t < 0) ? 1 : 0;
    }
    int64_t t = (int64_t)u[j + n] - borrow - (int64_t)carry;
    u[j + n] = (uint32_t)t;
    if (t < 0) { /* add back */
      uint64_t c = 0;
      for (int i = 0; i < n; i++) {
        uint64_t x = (uint64_t)u[i + j] + v[i] + c;
        u[i + j] = (uint32_t)x; c = x >> 32;
      }
      u[j + n] += (uint32_t)c;
    }
  }
  memset(r, 0, sizeof *r);
  for (int i = 0; i < n; i++) r->d[i] = s ? (u[i] >> s) | ((uint64_t)u[i + 1] << (32 - s)) : u[i];
  r->n = n; bn_trim(r);
}
static void bn_modexp(bn_t *r, const bn_t *base, uint64_t e, const bn_t *m) {
  bn_t acc, b, t;
  memset(&acc, 0, sizeof acc);
  acc.d[0] = 1; acc.n = 1;
  bn_mod(&acc, &acc, m); /* handles m == 1 */
  bn_mod(&b, base, m);
  int top = 63;
  while (top >= 0 && !((e >> top) & 1)) top--;
  for (int i = top; i >= 0; i--) {
    bn_mul(&t, &acc, &acc); bn_mod(&acc, &t, m);
    if ((e >> i) & 1) { bn_mul(&t, &acc, &b); bn_mod(&acc, &t, m); }
  }
  *r = acc;
}
static void bn_to_be(const bn_t *a, uint8_t *out, size_t len) {
  for (size_t i = 0; i < len; i++) {
    size_t bi = len - 1 - i;
    out[i] = (bi / 4 < (size_t)a->n) ? (uint8_t)(a->d[bi / 4] >> (8 * (bi % 4))) : 0;
  }
}
int zo_modexp(const uint8_t *base, size_t blen, uint64_t e, const uint8_t *mod, size_t mlen,
              uint8_t *out) {
  if (mlen > 4 * BN_MAX || blen > 4 * BN_MAX) return -1;
  bn_t b, m, r;
  bn_from_be(&b, base, blen);
  bn_from_be(&m, mod, mlen);
  if (m.n == 0) return -1;
#ifdef ZO_WITH_OPENSSL
  if (g_use_openssl) {
    BN_CTX *ctx = BN_CTX_new();
    BIGNUM *B = BN_bin2bn(base, (int)blen, NULL), *M = BN_bin2bn(mod, (int)mlen, NULL);
    BIGNUM *E = BN_new(), *R = BN_new();
    uint8_t eb[8];
    for (int i = 0; i < 8; i++) eb[i] = (uint8_t)(e >> (8 * (7 - i)));
    BN_bin2bn(eb, 8, E);
    BN_mod_exp(R, B, E, M, ctx);
    BN_bn2binpad(R, out, (int)mlen);
    BN_free(B); BN_free(M); BN_free(E); BN_free(R); BN_CTX_free(ctx);
    return 0;
  }
#endif
  bn_modexp(&r, &b, e, &m);
  bn_to_be(&r, out, mlen);
  return 0;
}

/* ------------------------------------------------------------------ PKCS#1 RSAPublicKey DER
 * rsa 0.9.6 RsaPublicKey::from_pkcs1_der (strict DER via the `der` crate) + check_public:
 * n odd, bits(n) <= 4096, 2 <= e <= 2^33-1, e odd, e < n  (SURVEY.md A.2 "RSA verify"). */
static int der_len(const uint8_t *p, size_t n, size_t *pos, size_t *out) {
  if (*pos >= n) return -1;
  uint8_t b = p[(*pos)++];
  if (b < 0x80) { *out = b; return 0; }
  int k = b & 0x7f;
  if (k == 0 || k > 4 || *pos + k > n) return -1;
  size_t v = 0;
  for (int i = 0; i < k; i++) v = (v << 8) | p[(*pos)++];
  if (v < 0x80) return -1;                       /* must have used short form */
  if (k > 1 && (v >> (8 * (k - 1))) == 0) return -1; /* non-minimal */
  *out = v;
  return 0;
}
static int der_uint(const uint8_t *p, size_t n, size_t *pos, const uint8_t **val, size_t *vlen) {
  if (*pos >= n || p[(*pos)++] != 0x02) return -1;
  size_t l;
  if (der_len(p, n, pos, &l) || l == 0 || *pos + l > n) return -1;
  const uint8_t *v = p + *pos;
  if (v[0] & 0x80) return -1;                          /* negative */
  if (l > 1 && v[0] == 0 && !(v[1] & 0x80)) return -1; /* non-minimal */
  if (l > 1 && v[0] == 0) { v++; l--; }
  *val = v; *vlen = l; /* minimal magnitude; zero is a single 0x00 */
  *pos = (size_t)(v - p) + l;
  return 0;
}
int zo_parse_rsa_der(const uint8_t *der, size_t len, uint8_t *n_out, size_t *n_len, uint64_t *e) {
  size_t pos = 0, sl;
  if (len < 2 || der[pos++] != 0x30) return -1;
  if (der_len(der, len, &pos, &sl)) return -1;
  if (pos + sl != len) return -1; /* trailing data / truncated */
  const uint8_t *nv, *ev;
  size_t nl, el;
  if (der_uint(der, len, &pos, &nv, &nl)) return -1;
  if (der_uint(der, len, &pos, &ev, &el)) return -1;
  if (pos != len) return -1;
  /* check_public_with_max_size */
  bn_t N;
  if (nl > 4 * BN_MAX) return -2;
  bn_from_be(&N, nv, nl);
  if (bn_bits(&N) > 4096) return -2;
  if (el > 8) return -3; /* e.to_u64() fails */
  uint64_t E = 0;
  for (size_t i = 0; i < el; i++) E = (E << 8) | ev[i];
  bn_t Eb;
  memset(&Eb, 0, sizeof Eb);
  Eb.d[0] = (uint32_t)E; Eb.d[1] = (uint32_t)(E >> 32); Eb.n = 2; bn_trim(&Eb);
  if (bn_cmp(&Eb, &N) >= 0 || N.n == 0 || !(N.d[0] & 1)) return -4;
  if (!(E & 1)) return -5;
  if (E < 2) return -6;
  if (E > ((1ull << 33) - 1)) return -7;
  if (n_out) memcpy(n_out, nv, nl);
  if (n_len) *n_len = nl;
  if (e) *e = E;
  return 0;
}

/* rsa 0.9.6 pkcs1v15::verify + pkcs1v15_sign_unpad with the SHA-256 DigestInfo prefix */
static const uint8_t SHA256_PREFIX[19] = {0x30, 0x31, 0x30, 0x0d, 0x06, 0x09, 0x60, 0x86, 0x48, 0x01,
                                          0x65, 0x03, 0x04, 0x02, 0x01, 0x05, 0x00, 0x04, 0x20};
int zo_rsa_verify_sha256(const uint8_t *der, size_t der_len_, const uint8_t hash[32],
                         const uint8_t *sig, size_t sig_len) {
  uint8_t nbuf[4 * BN_MAX];
  size_t nl;
  uint64_t e;
  if (zo_parse_rsa_der(der, der_len_, nbuf, &nl, &e)) return -1;
  bn_t N, S;
  bn_from_be(&N, nbuf, nl);
  size_t k = (size_t)(bn_bits(&N) + 7) / 8;
  if (sig_len > 4 * BN_MAX) return 0;
  bn_from_be(&S, sig, sig_len);
  if (bn_cmp(&S, &N) >= 0 || sig_len != k) return 0;
  uint8_t em[4 * BN_MAX];
  if (zo_modexp(sig, sig_len, e, nbuf, nl, em + (k - nl))) return 0;
  memset(em, 0, k - nl);
  size_t t_len = 19 + 32;
  if (k < t_len + 11) return 0;
  int ok = (em[0] == 0) & (em[1] == 1);
  ok &= memcmp(em + k - 32, hash, 32) == 0;
  ok &= memcmp(em + k - t_len, SHA256_PREFIX, 19) == 0;
  ok &= em[k - t_len - 1] == 0;
  for (size_t i = 2; i < k - t_len - 1; i++) ok &= em[i] == 0xff;
  return ok;
}

/* ------------------------------------------------------------------ String::from_utf8_lossy */
size_t zo_utf8_lossy(const uint8_t *in, size_t n, uint8_t *out) {
  size_t i = 0, o = 0;
#define GETB(k) ((i + (k)) < n ? in[i + (k)] : 0)
  while (i < n) {
    uint8_t b = in[i];
    size_t bad = 0, w = 0;
    if (b < 0x80) w = 1;
    else if (b >= 0xC2 && b <= 0xDF) {
      if ((GETB(1) & 0xC0) == 0x80) w = 2; else bad = 1;
    } else if (b >= 0xE0 && b <= 0xEF) {
      uint8_t c = GETB(1);
      int ok2 = (b == 0xE0) ? (c >= 0xA0 && c <= 0xBF)
              : (b == 0xED) ? (c >= 0x80 && c <= 0x9F) : (c >= 0x80 && c <= 0xBF);
      if (!ok2) bad = 1;
      else if ((GETB(2) & 0xC0) != 0x80) bad = 2;
      else w = 3;
    } else if (b >= 0xF0 && b <= 0xF4) {
      uint8_t c = GETB(1);
      int ok2 = (b == 0xF0) ? (c >= 0x90 && c <= 0xBF)
              : (b == 0xF4) ? (c >= 0x80 && c <= 0x8F) : (c >= 0x80 && c <= 0xBF);
      if (!ok2) bad = 1;
      else if ((GETB(2) & 0xC0) != 0x80) bad = 2;
      else if ((GETB(3) & 0xC0) != 0x80) bad = 3;
      else w = 4;
    } else bad = 1;
    if (w) { memcpy(out + o, in + i, w); o += w; i += w; }
    else { out[o++] = 0xEF; out[o++] = 0xBF; out[o++] = 0xBD; i += bad; }
  }
#undef GETB
  return o;
}

/* ------------------------------------------------------------------ mailparse 0.15.0
 * parse_header / parse_headers state machines ([dep-memory], SURVEY.md A.2 "mailparse header
 * split").  Only the failure modes that make parse_mail() return Err at the top level are
 * modelled; MIME sub-part recursion is NOT (documented gap). */
static long mp_parse_header(const uint8_t *d, size_t n, size_t *key_end, int *has_key,
                            size_t *vs, size_t *ve) {
  enum { S_INIT, S_KEY, S_PREV, S_VAL, S_VALNL } st = S_INIT;
  size_t ix = 0, ix_key_end = 0, vstart = 0, vend = 0;
  int hk = 0;
  if (n == 0) return -1;
  uint8_t c = d[0];
  for (;;) {
    switch (st) {
      case S_INIT:
        if (c == ' ') return -1;
        st = S_KEY;
        continue;
      case S_KEY:
        if (c == ':') { ix_key_end = ix; hk = 1; st = S_PREV; }
        else if (c == '\n') { vstart = ix; vend = ix; ix++; goto done; }
        break;
      case S_PREV:
        if (c != ' ') { vstart = ix; vend = ix; st = S_VAL; continue; }
        break;
      case S_VAL:
        if (c == '\n') st = S_VALNL;
        else if (c != '\r') vend = ix + 1;
        break;
      case S_VALNL:
        if (c == ' ' || c == '\t') { st = S_VAL; continue; }
        goto done;
    }
    ix++;
    if (ix >= n) break;
    c = d[ix];
  }
done:
  *has_key = hk;
  *key_end = hk ? ix_key_end : vstart;
  *vs = vstart; *ve = vend;
  return (long)ix;
}
long zo_parse_headers(const uint8_t *raw, size_t n, uint32_t *quads, size_t cap, size_t *body_off) {
  size_t ix = 0;
  long cnt = 0;
  for (;;) {
    if (ix >= n) break;
    if (raw[ix] == '\n') { ix += 1; break; }
    if (raw[ix] == '\r') {
      if (ix + 1 < n && raw[ix + 1] == '\n') { ix += 2; break; }
      return -1; /* lone CR after headers */
    }
    size_t ke, vs, ve;
    int hk;
    long used = mp_parse_header(raw + ix, n - ix, &ke, &hk, &vs, &ve);
    if (used < 0) return -1;
    if ((size_t)cnt < cap) {
      quads[4 * cnt + 0] = (uint32_t)ix;
      quads[4 * cnt + 1] = (uint32_t)ke;
      quads[4 * cnt + 2] = (uint32_t)(ix + vs);
      quads[4 * cnt + 3] = (uint32_t)(ve - vs);
    }
    cnt++;
    ix += (size_t)used;
  }
  if (body_off) *body_off = ix;
  return cnt;
}

/* ------------------------------------------------------------------ cfdkim canonicalization.rs */
size_t zo_canon_body_relaxed(const uint8_t *in, size_t n, uint8_t *out) {
  /* (1) TAB->SP, (2) collapse SP runs, (3) delete the SP of every " \r\n" (repeated scan, as the
   * reference does), (4) strip trailing empty lines, (5) append CRLF if non-empty w/o one. */
  size_t m = 0;
  int prev = 0;
  for (size_t i = 0; i < n; i++) {
    uint8_t c = in[i] == '\t' ? ' ' : in[i];
    if (c == ' ') { if (prev) continue; prev = 1; } else prev = 0;
    out[m++] = c;
  }
  for (;;) { /* while let Some(idx) = find(" \r\n") { remove(idx) } */
    size_t idx = (size_t)-1;
    for (size_t i = 0; i + 3 <= m; i++)
      if (out[i] == ' ' && out[i + 1] == '\r' && out[i + 2] == '\n') { idx = i; break; }
    if (idx == (size_t)-1) break;
    memmove(out + idx, out + idx + 1, m - idx - 1);
    m--;
  }
  while (m >= 4 && memcmp(out + m - 4, "\r\n\r\n", 4) == 0) m -= 2;
  if (m > 0 && !(m >= 2 && out[m - 2] == '\r' && out[m - 1] == '\n')) { out[m++] = '\r'; out[m++] = '\n'; }
  return m;
}
size_t zo_canon_body_simple(const uint8_t *in, size_t n, uint8_t *out) {
  if (n == 0) { out[0] = '\r'; out[1] = '\n'; return 2; }
  while (n >= 4 && memcmp(in + n - 4, "\r\n\r\n", 4) == 0) n -= 2;
  memcpy(out, in, n);
  return n;
}
/* mailparse get_key(): bytes decoded as Latin-1 into a String (so >=0x80 becomes 2 UTF-8 bytes). */
static size_t latin1_to_utf8(const uint8_t *k, size_t n, uint8_t *out, int lower) {
  size_t o = 0;
  for (size_t i = 0; i < n; i++) {
    uint32_t c = k[i];
    if (lower) { /* char::to_lowercase restricted to U+0000..U+00FF */
      if (c >= 'A' && c <= 'Z') c += 32;
      else if (c >= 0xC0 && c <= 0xDE && c != 0xD7) c += 32;
    }
    if (c < 0x80) out[o++] = (uint8_t)c;
    else { out[o++] = (uint8_t)(0xC0 | (c >> 6)); out[o++] = (uint8_t)(0x80 | (c & 0x3F)); }
  }
  return o;
}
static int latin1_is_ws(uint8_t c) { /* char::is_whitespace within Latin-1 */
  return c == ' ' || (c >= 9 && c <= 13) || c == 0x85 || c == 0xA0;
}
size_t zo_canon_header_relaxed(const uint8_t *key, size_t klen, const uint8_t *val, size_t vlen,
                               uint8_t *out) {
  while (klen > 0 && latin1_is_ws(key[klen - 1])) klen--; /* key.to_lowercase().trim_end() */
  size_t o = latin1_to_utf8(key, klen, out, 1);
  out[o++] = ':';
  /* value: TAB->SP; remove every "\r\n"; pop trailing SP; strip leading SP; collapse SP runs */
  uint8_t *v = out + o;
  size_t m = 0;
  for (size_t i = 0; i < vlen; i++) {
    uint8_t c = val[i] == '\t' ? ' ' : val[i];
    if (c == '\r' && i + 1 < vlen && val[i + 1] == '\n') { i++; continue; }
    v[m++] = c;
  }
  while (m > 0 && v[m - 1] == ' ') m--;
  size_t lead = 0;
  while (lead < m && v[lead] == ' ') lead++;
  size_t w = 0;
  int prev = 0;
  for (size_t i = lead; i < m; i++) {
    uint8_t c = v[i];
    if (c == ' ') { if (prev) continue; prev = 1; } else prev = 0;
    v[w++] = c;
  }
  o += w;
  out[o++] = '\r'; out[o++] = '\n';
  return o;
}
size_t zo_canon_header_simple(const uint8_t *key, size_t klen, const uint8_t *val, size_t vlen,
                              uint8_t *out) {
  size_t o = latin1_to_utf8(key, klen, out, 0);
  out[o++] = ':'; out[o++] = ' ';
  memcpy(out + o, val, vlen); o += vlen;
  out[o++] = '\r'; out[o++] = '\n';
  return o;
}

/* ------------------------------------------------------------------ cfdkim parser.rs tag-list */
typedef struct {
  size_t name_off, name_len;
  uint8_t *value; size_t value_len;   /* FWS removed */
  size_t raw_off, raw_len;            /* slice of the header value string, FWS kept */
} tag_t;
#define MAX_TAGS 64
typedef struct {
  const uint8_t *s; size_t n; /* the (lossy utf-8) header value string */
  tag_t tags[MAX_TAGS];
  int n_tags;
} dkim_header_t;

static int is_fws(uint8_t c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n'; }
static int is_valchar(uint8_t c) { return (c >= 0x21 && c <= 0x3A) || (c >= 0x3C && c <= 0x7E); }
static int is_alpha(uint8_t c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'); }
static int is_alnum_(uint8_t c) { return is_alpha(c) || (c >= '0' && c <= '9') || c == '_'; }

/* tag-spec = [FWS] tag-name [FWS] "=" [FWS] tag-value [FWS]; returns new pos or -1 */
static long parse_tag_spec(const uint8_t *s, size_t n, size_t pos, tag_t *t) {
  while (pos < n && is_fws(s[pos])) pos++;
  if (pos >= n || !is_alpha(s[pos])) return -1;
  size_t ns = pos;
  while (pos < n && is_alnum_(s[pos])) pos++;
  t->name_off = ns; t->name_len = pos - ns;
  while (pos < n && is_fws(s[pos])) pos++;
  if (pos >= n || s[pos] != '=') return -1;
  pos++;
  while (pos < n && is_fws(s[pos])) pos++;
  buf_t v = {0};
  buf_reserve(&v, 1);
  size_t raw_s = pos, raw_e = pos;
  if (pos < n && is_valchar(s[pos])) {
    size_t a = pos;
    while (pos < n && is_valchar(s[pos])) pos++;
    buf_put(&v, s + a, pos - a);
    raw_e = pos;
    for (;;) { /* *( FWS tval ) */
      size_t q = pos;
      while (q < n && is_fws(s[q])) q++;
      if (q == pos || q >= n || !is_valchar(s[q])) break;
      size_t b = q;
      while (q < n && is_valchar(s[q])) q++;
      buf_put(&v, s + b, q - b);
      pos = q; raw_e = q;
    }
  }
  while (pos < n && is_fws(s[pos])) pos++;
  t->value = v.p; t->value_len = v.n;
  t->raw_off = raw_s; t->raw_len = raw_e - raw_s;
  return (long)pos;
}
static void dkim_header_free(dkim_header_t *h) {
  for (int i = 0; i < h->n_tags; i++) free(h->tags[i].value);
  h->n_tags = 0;
}
static tag_t *get_tag(dkim_header_t *h, const char *name) {
  size_t l = strlen(name);
  for (int i = 0; i < h->n_tags; i++)
    if (h->tags[i].name_len == l && memcmp(h->s + h->tags[i].name_off, name, l) == 0) return &h->tags[i];
  return NULL;
}
static int tag_eq(const tag_t *t, const char *lit) {
  size_t l = strlen(lit);
  return t->value_len == l && memcmp(t->value, lit, l) == 0;
}
/* cfdkim::validate_header (lib.rs); returns ZO_DKIM_PASS or an error kind */
static int validate_header(const uint8_t *s, size_t n, int64_t now_unix, dkim_header_t *h) {
  h->s = s; h->n = n; h->n_tags = 0;
  tag_t t;
  long pos = parse_tag_spec(s, n, 0, &t);
  if (pos < 0) return ZO_DKIM_SYNTAX;
  int seen_overflow = 0;
  for (;;) {
    /* IndexMap insert: later duplicates overwrite the value, keep the first position */
    int dup = -1;
    for (int i = 0; i < h->n_tags; i++)
      if (h->tags[i].name_len == t.name_len &&
          memcmp(s + h->tags[i].name_off, s + t.name_off, t.name_len) == 0) dup = i;
    if (dup >= 0) { free(h->tags[dup].value); h->tags[dup] = t; }
    else if (h->n_tags < MAX_TAGS) h->tags[h->n_tags++] = t;
    else { free(t.value); seen_overflow = 1; }
    if ((size_t)pos >= n || s[pos] != ';') break;
    long np = parse_tag_spec(s, n, (size_t)pos + 1, &t);
    if (np < 0) break; /* trailing ";" or garbage: remaining input ignored */
    pos = np;
  }
  (void)seen_overflow;
  static const char *REQ[] = {"v", "a", "b", "bh", "d", "h", "s"};
  for (int i = 0; i < 7; i++)
    if (!get_tag(h, REQ[i])) return ZO_DKIM_MISSING_TAG;
  if (!tag_eq(get_tag(h, "v"), "1")) return ZO_DKIM_VERSION;
  tag_t *ti = get_tag(h, "i"), *td = get_tag(h, "d");
  if (ti) { /* user.ends_with(signing_domain) */
    if (ti->value_len < td->value_len ||
        memcmp(ti->value + ti->value_len - td->value_len, td->value, td->value_len) != 0)
      return ZO_DKIM_DOMAIN_MISMATCH;
  }
  { /* h= split(':') lower-cased must contain "from" */
    tag_t *th = get_tag(h, "h");
    int found = 0;
    size_t a = 0;
    for (size_t i = 0; i <= th->value_len; i++) {
      if (i == th->value_len || th->value[i] == ':') {
        if (i - a == 4) {
          uint8_t w[4];
          for (int k = 0; k < 4; k++) { uint8_t c = th->value[a + k]; w[k] = (c >= 'A' && c <= 'Z') ? c + 32 : c; }
          if (memcmp(w, "from", 4) == 0) found = 1;
        }
        a = i + 1;
      }
    }
    if (!found) return ZO_DKIM_FROM_NOT_SIGNED;
  }
  tag_t *tq = get_tag(h, "q");
  if (tq && !tag_eq(tq, "dns/txt")) return ZO_DKIM_QUERY_METHOD;
  tag_t *tx = get_tag(h, "x");
  if (tx) { /* x.parse::<i64>().unwrap_or_default(); now > x + 15 min => expired */
    int64_t x = 0;
    int ok = tx->value_len > 0;
    size_t i = 0;
    int neg = 0;
    if (ok && (tx->value[0] == '+' || tx->value[0] == '-')) { neg = tx->value[0] == '-'; i = 1; ok = tx->value_len > 1; }
    for (; ok && i < tx->value_len; i++) {
      uint8_t c = tx->value[i];
      if (c < '0' || c > '9') { ok = 0; break; }
      int d = c - '0';
      if (!neg) { if (x > (INT64_MAX - d) / 10) { ok = 0; break; } x = x * 10 + d; }
      else { if (x < (INT64_MIN + d) / 10) { ok = 0; break; } x = x * 10 - d; }
    }
    if (!ok) x = 0;
    int64_t now = now_unix ? now_unix : (int64_t)time(NULL);
    if (now > x + 15 * 60) return ZO_DKIM_EXPIRED;
  }
  return ZO_DKIM_PASS;
}

/* ------------------------------------------------------------------ header selection + hashing */
typedef struct {
  const uint8_t *raw; size_t n;
  uint32_t *q; long cnt; /* header quads */
} mail_t;
static int ascii_ieq(const uint8_t *a, size_t al, const uint8_t *b, size_t bl) {
  if (al != bl) return 0;
  for (size_t i = 0; i < al; i++) {
    uint8_t x = a[i], y = b[i];
    if (x >= 'A' && x <= 'Z') x += 32;
    if (y >= 'A' && y <= 'Z') y += 32;
    if (x != y) return 0;
  }
  return 1;
}
/* cfdkim hash.rs select_headers + compute_headers_hash preimage ([dep-memory]) */
static void build_header_preimage(const mail_t *m, dkim_header_t *dh, int relaxed, buf_t *out) {
  tag_t *th = get_tag(dh, "h");
  /* per-name cursor, kept in a small assoc list */
  struct { size_t off, len; long idx; } last[128];
  int n_last = 0;
  size_t a = 0;
  for (size_t i = 0; i <= th->value_len; i++) {
    if (i != th->value_len && th->value[i] != ':') continue;
    size_t s = a, e = i;
    a = i + 1;
    while (s < e && latin1_is_ws(th->value[s]) && th->value[s] < 0x80) s++;
    while (e > s && latin1_is_ws(th->value[e - 1]) && th->value[e - 1] < 0x80) e--;
    if (s == e) continue;
    long start = m->cnt;
    int li = -1;
    for (int k = 0; k < n_last; k++)
      if (ascii_ieq(th->value + last[k].off, last[k].len, th->value + s, e - s)) { li = k; start = last[k].idx; }
    long hit = -1;
    for (long j = start - 1; j >= 0; j--) {
      const uint32_t *q = m->q + 4 * j;
      /* key compare: decode_latin1(key).eq_ignore_ascii_case(name): equal only for ASCII keys */
      if (ascii_ieq(m->raw + q[0], q[1], th->value + s, e - s)) { hit = j; break; }
    }
    if (li < 0 && n_last < 128) { li = n_last++; last[li].off = s; last[li].len = e - s; }
    if (li >= 0) last[li].idx = hit >= 0 ? hit : 0;
    if (hit >= 0) {
      const uint32_t *q = m->q + 4 * hit;
      buf_reserve(out, 2 * q[1] + q[3] + 8);
      out->n += relaxed ? zo_canon_header_relaxed(m->raw + q[0], q[1], m->raw + q[2], q[3], out->p + out->n)
                        : zo_canon_header_simple(m->raw + q[0], q[1], m->raw + q[2], q[3], out->p + out->n);
    }
  }
  /* DKIM-Signature header itself: value.replace(raw_b, "") canonicalised, final CRLF removed */
  tag_t *tb = get_tag(dh, "b");
  buf_t v = {0};
  buf_reserve(&v, dh->n + 1);
  if (tb->raw_len == 0) buf_put(&v, dh->s, dh->n);
  else {
    size_t i = 0;
    const uint8_t *pat = dh->s + tb->raw_off;
    while (i < dh->n) {
      if (i + tb->raw_len <= dh->n && memcmp(dh->s + i, pat, tb->raw_len) == 0) i += tb->raw_len;
      else buf_putc(&v, dh->s[i++]);
    }
  }
  buf_reserve(out, v.n + 64);
  size_t w = relaxed ? zo_canon_header_relaxed((const uint8_t *)"DKIM-Signature", 14, v.p, v.n, out->p + out->n)
                     : zo_canon_header_simple((const uint8_t *)"DKIM-Signature", 14, v.p, v.n, out->p + out->n);
  out->n += w - 2;
  buf_free(&v);
}
static const uint8_t *find_body(const uint8_t *raw, size_t n, size_t *blen) {
  /* bytes::get_all_after(raw, "\r\n\r\n") */
  for (size_t i = 0; i + 4 <= n; i++)
    if (raw[i] == '\r' && raw[i + 1] == '\n' && raw[i + 2] == '\r' && raw[i + 3] == '\n') {
      *blen = n - i - 4;
      return raw + i + 4;
    }
  *blen = 0;
  return raw + n;
}
static int parse_canon(dkim_header_t *dh, int *hdr_relaxed, int *body_relaxed) {
  tag_t *tc = get_tag(dh, "c");
  if (!tc) { *hdr_relaxed = 0; *body_relaxed = 0; return 0; }
  if (tag_eq(tc, "simple/simple")) { *hdr_relaxed = 0; *body_relaxed = 0; }
  else if (tag_eq(tc, "relaxed/simple")) { *hdr_relaxed = 1; *body_relaxed = 0; }
  else if (tag_eq(tc, "simple/relaxed")) { *hdr_relaxed = 0; *body_relaxed = 1; }
  else if (tag_eq(tc, "relaxed/relaxed")) { *hdr_relaxed = 1; *body_relaxed = 1; }
  else if (tag_eq(tc, "relaxed")) { *hdr_relaxed = 1; *body_relaxed = 0; }
  else if (tag_eq(tc, "simple")) { *hdr_relaxed = 0; *body_relaxed = 0; }
  else return -1;
  return 0;
}
/* usize::from_str: optional leading '+', digits, overflow => Err */
static int parse_usize(const tag_t *t, size_t *out) {
  size_t i = 0;
  if (t->value_len && t->value[0] == '+') i = 1;
  if (i >= t->value_len) return -1;
  uint64_t v = 0;
  for (; i < t->value_len; i++) {
    uint8_t c = t->value[i];
    if (c < '0' || c > '9') return -1;
    if (v > (UINT64_MAX - (c - '0')) / 10) return -1;
    v = v * 10 + (c - '0');
  }
  *out = (size_t)v;
  return 0;
}
static int canon_body_for(const mail_t *m, dkim_header_t *dh, int body_relaxed, buf_t *out) {
  size_t bl;
  const uint8_t *b = find_body(m->raw, m->n, &bl);
  buf_reserve(out, bl + 4);
  out->n = body_relaxed ? zo_canon_body_relaxed(b, bl, out->p) : zo_canon_body_simple(b, bl, out->p);
  tag_t *tl = get_tag(dh, "l");
  if (tl) {
    size_t l;
    if (parse_usize(tl, &l)) return -1;
    if (l < out->n) out->n = l; /* Vec::truncate */
  }
  return 0;
}

static int mail_parse(const uint8_t *raw, size_t n, mail_t *m) {
  m->raw = raw; m->n = n;
  long cnt = zo_parse_headers(raw, n, NULL, 0, NULL);
  if (cnt < 0) return -1;
  m->q = (uint32_t *)malloc(sizeof(uint32_t) * 4 * (size_t)(cnt ? cnt : 1));
  m->cnt = zo_parse_headers(raw, n, m->q, (size_t)cnt, NULL);
  return 0;
}

/* cfdkim::verify_email_with_key + verify_email_header (SURVEY.md A.2 pseudo-code) */
static void verify_dkim(const zo_email *e, const mail_t *m, const uint8_t *rsa_der, size_t der_len_,
                        int64_t now_unix, zo_result *out) {
  int have_err = 0, last_err = ZO_DKIM_NEUTRAL, saw_unsupported = 0;
  for (long hi = 0; hi < m->cnt; hi++) {
    const uint32_t *q = m->q + 4 * hi;
    if (!ascii_ieq(m->raw + q[0], q[1], (const uint8_t *)"DKIM-Signature", 14)) continue;
    uint8_t *val = (uint8_t *)malloc(3 * (size_t)q[3] + 1);
    size_t vl = zo_utf8_lossy(m->raw + q[2], q[3], val);
    dkim_header_t dh;
    int r = validate_header(val, vl, now_unix, &dh);
    if (r != ZO_DKIM_PASS) { have_err = 1; last_err = r; goto next; }
    {
      tag_t *td = get_tag(&dh, "d");
      /* signing_domain.to_lowercase() != from_domain.to_lowercase(): d= is ASCII (VALCHAR);
       * from_domain lower-cased ASCII-only here (non-ASCII from_domain can never equal d=) */
      if (!ascii_ieq(td->value, td->value_len, (const uint8_t *)e->from_domain, e->from_domain_len)) goto next;
      int hr, br;
      if (parse_canon(&dh, &hr, &br)) { have_err = 1; last_err = ZO_DKIM_CANON_TYPE; goto next; }
      tag_t *ta = get_tag(&dh, "a");
      int algo; /* 0 rsa-sha1, 1 rsa-sha256, 2 ed25519-sha256 */
      if (tag_eq(ta, "rsa-sha1")) algo = 0;
      else if (tag_eq(ta, "rsa-sha256")) algo = 1;
      else if (tag_eq(ta, "ed25519-sha256")) algo = 2;
      else { have_err = 1; last_err = ZO_DKIM_HASH_ALGO; goto next; }
      if (algo == 0) { /* rsa-sha1: the reference would verify with SHA-1; outcome unknown to us.
                        * Keep scanning: a later passing signature decides "pass" either way. */
        saw_unsupported = 1; goto next;
      }
      buf_t body = {0}, pre = {0};
      if (canon_body_for(m, &dh, br, &body)) { have_err = 1; last_err = ZO_DKIM_LENGTH_TAG; buf_free(&body); goto next; }
      build_header_preimage(m, &dh, hr, &pre);
      zo_sha256(body.p, body.n, out->body_hash);
      zo_sha256(pre.p, pre.n, out->header_hash);
      buf_free(&body); buf_free(&pre);
      char b64[48];
      size_t bl = zo_base64_encode(out->body_hash, 32, b64);
      tag_t *tbh = get_tag(&dh, "bh");
      out->bh_ok = (tbh->value_len == bl && memcmp(tbh->value, b64, bl) == 0);
      out->rsa_ok = 0;
      if (!out->bh_ok) { have_err = 1; last_err = ZO_DKIM_BODY_HASH; goto next; }
      tag_t *tb = get_tag(&dh, "b");
      uint8_t *sig = (uint8_t *)malloc(tb->value_len + 4);
      long sl = zo_base64_decode(tb->value, tb->value_len, sig);
      if (sl < 0) { free(sig); have_err = 1; last_err = ZO_DKIM_SIG_SYNTAX; goto next; }
      if (algo == 2) { free(sig); have_err = 1; last_err = ZO_DKIM_ALGO_KEY_MISMATCH; goto next; }
      int ok = zo_rsa_verify_sha256(rsa_der, der_len_, out->header_hash, sig, (size_t)sl);
      free(sig);
      if (ok == 1) {
        out->rsa_ok = 1; out->dkim_detail = ZO_DKIM_PASS;
        dkim_header_free(&dh); free(val);
        return;
      }
      have_err = 1; last_err = ZO_DKIM_SIG_MISMATCH;
    }
  next:
    dkim_header_free(&dh);
    free(val);
  }
  out->dkim_detail = have_err ? last_err : ZO_DKIM_NEUTRAL;
  if (saw_unsupported) out->status = ZO_ERR_UNSUPPORTED;
}

/* cfdkim::canonicalize_signed_email ([dep-memory]; SURVEY.md A.2 last bullet) */
int zo_canonicalize_signed_email(const uint8_t *raw, size_t n, int64_t now_unix, uint8_t **hdr,
                                 size_t *hdr_len, uint8_t **body, size_t *body_len) {
  mail_t m;
  if (mail_parse(raw, n, &m)) return 1;
  int rc = 2; /* no valid DKIM-Signature header */
  for (long hi = 0; hi < m.cnt; hi++) {
    const uint32_t *q = m.q + 4 * hi;
    if (!ascii_ieq(m.raw + q[0], q[1], (const uint8_t *)"DKIM-Signature", 14)) continue;
    uint8_t *val = (uint8_t *)malloc(3 * (size_t)q[3] + 1);
    size_t vl = zo_utf8_lossy(m.raw + q[2], q[3], val);
    dkim_header_t dh;
    int r = validate_header(val, vl, now_unix, &dh);
    if (r != ZO_DKIM_PASS) { dkim_header_free(&dh); free(val); continue; }
    int hr, br;
    buf_t b = {0}, p = {0};
    if (parse_canon(&dh, &hr, &br)) rc = 3;
    else if (canon_body_for(&m, &dh, br, &b)) rc = 4;
    else {
      tag_t *tb = get_tag(&dh, "b");
      uint8_t *sig = (uint8_t *)malloc(tb->value_len + 4);
      long sl = zo_base64_decode(tb->value, tb->value_len, sig);
      free(sig);
      if (sl < 0) rc = 5;
      else {
        build_header_preimage(&m, &dh, hr, &p);
        buf_reserve(&b, 1); buf_reserve(&p, 1);
        *hdr = p.p; *hdr_len = p.n; *body = b.p; *body_len = b.n;
        p.p = NULL; b.p = NULL;
        rc = 0;
      }
    }
    buf_free(&b); buf_free(&p);
    dkim_header_free(&dh); free(val);
    break;
  }
  free(m.q);
  return rc;
}
void zo_free(void *p) { free(p); }

/* ------------------------------------------------------------------ core/src/email.rs:61-86 */
size_t zo_qp_clean(const uint8_t *in, size_t n, uint8_t *out) {
  size_t o = 0;
  for (size_t i = 0; i < n;) {
    if (in[i] == '=' && i + 3 <= n && in[i + 1] == '\r' && in[i + 2] == '\n') i += 3;
    else out[o++] = in[i++];
  }
  size_t cleaned = o;
  while (o < n) out[o++] = 0;
  return cleaned;
}

/* ------------------------------------------------------------------ ZDF1 dense DFA tables and
 * regex-automata 0.4.9 dfa::regex::Regex search semantics (SURVEY.md A.5 pseudo-code) */
typedef struct {
  uint32_t flags, n_states, n_classes, min_match, max_match;
  uint32_t start[12];
  const uint8_t *class_map, *start_map;
  const uint8_t *trans; /* u32 LE, unaligned-safe access below */
} zdf_t;
#define ZDF_MAGIC 0x3146445Au
#define ZDF_REVERSE 1u
#define ZDF_UTF8 2u
#define ZDF_HAS_EMPTY 4u
static uint32_t rd32(const uint8_t *p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static int zdf_load(const uint8_t *b, size_t n, zdf_t *d) {
  if (n < 584 || rd32(b) != ZDF_MAGIC) return -1;
  d->flags = rd32(b + 4); d->n_states = rd32(b + 8); d->n_classes = rd32(b + 12);
  d->min_match = rd32(b + 16); d->max_match = rd32(b + 20);
  for (int i = 0; i < 12; i++) d->start[i] = rd32(b + 24 + 4 * i);
  d->class_map = b + 72; d->start_map = b + 328; d->trans = b + 584;
  if (d->n_states == 0 || d->n_classes < 2 || d->n_classes > 257) return -1;
  if ((uint64_t)d->n_states * d->n_classes * 4 + 584 != n) return -1;
  for (int i = 0; i < 12; i++) if (d->start[i] >= d->n_states) return -1;
  for (int i = 0; i < 256; i++) {
    if (d->class_map[i] >= d->n_classes - 1) return -1;
    if (d->start_map[i] > 5) return -1;
  }
  for (uint64_t i = 0; i < (uint64_t)d->n_states * d->n_classes; i++)
    if (rd32(d->trans + 4 * i) >= d->n_states) return -1;
  return 0;
}
static inline uint32_t zdf_next(const zdf_t *d, uint32_t s, uint32_t cls) {
  return rd32(d->trans + 4 * ((size_t)s * d->n_classes + cls));
}
static inline int zdf_is_match(const zdf_t *d, uint32_t s) { return s >= d->min_match && s <= d->max_match; }
static int is_char_boundary(const uint8_t *h, size_t n, size_t i) {
  if (i == n) return 1;
  if (i > n) return 0;
  return (int8_t)h[i] >= -0x40;
}
/* forward unanchored leftmost-first search in [start,n). returns 1 and *end on match */
static int find_fwd_raw(const zdf_t *d, const uint8_t *h, size_t n, size_t start, size_t *end) {
  if (start > n) return 0;
  uint32_t kind = start == 0 ? 2 : d->start_map[h[start - 1]];
  uint32_t sid = d->start[kind];
  int have = 0;
  if (sid == 0) return 0;
  for (size_t at = start; at < n; at++) {
    sid = zdf_next(d, sid, d->class_map[h[at]]);
    if (sid == 0) { if (have) return 1; return 0; }
    if (zdf_is_match(d, sid)) { have = 1; *end = at; }
  }
  sid = zdf_next(d, sid, d->n_classes - 1); /* EOI */
  if (zdf_is_match(d, sid)) { have = 1; *end = n; }
  return have;
}
static int find_fwd(const zdf_t *d, const uint8_t *h, size_t n, size_t start, size_t *end) {
  int r = find_fwd_raw(d, h, n, start, end);
  if (!r) return 0;
  if ((d->flags & ZDF_UTF8) && (d->flags & ZDF_HAS_EMPTY)) { /* util::empty::skip_splits_fwd */
    while (!is_char_boundary(h, n, *end)) {
      start++;
      if (!find_fwd_raw(d, h, n, start, end)) return 0;
    }
  }
  return 1;
}
/* anchored reverse search over [start,end): leftmost start of a match ending at `end` */
static int find_rev(const zdf_t *d, const uint8_t *h, size_t n, size_t start, size_t end, size_t *ms) {
  uint32_t kind = end == n ? 2 : d->start_map[h[end]];
  uint32_t sid = d->start[6 + kind];
  int have = 0;
  if (sid == 0) return 0;
  for (size_t at = end; at > start;) {
    at--;
    sid = zdf_next(d, sid, d->class_map[h[at]]);
    if (sid == 0) return have;
    if (zdf_is_match(d, sid)) { have = 1; *ms = at + 1; }
  }
  sid = start == 0 ? zdf_next(d, sid, d->n_classes - 1) : zdf_next(d, sid, d->class_map[h[start - 1]]);
  if (zdf_is_match(d, sid)) { have = 1; *ms = start; }
  return have;
}
static int regex_find(const zdf_t *f, const zdf_t *r, const uint8_t *h, size_t n, size_t start,
                      size_t *ms, size_t *me) {
  size_t e;
  if (!find_fwd(f, h, n, start, &e)) return 0;
  if (e == start) { *ms = e; *me = e; return 1; }
  size_t s;
  if (!find_rev(r, h, n, start, e, &s)) return -1; /* "reverse search must match" panic */
  *ms = s; *me = e;
  return 1;
}
long zo_dfa_find_iter(const uint8_t *fwd, size_t fwd_len, const uint8_t *bwd, size_t bwd_len,
                      const uint8_t *hay, size_t n, uint32_t *spans, size_t cap) {
  zdf_t f, r;
  if (zdf_load(fwd, fwd_len, &f) || zdf_load(bwd, bwd_len, &r)) return -1;
  if ((f.flags & ZDF_REVERSE) || !(r.flags & ZDF_REVERSE)) return -1;
  long cnt = 0;
  size_t pos = 0, last_end = (size_t)-1;
  for (;;) {
    size_t s, e;
    int k = regex_find(&f, &r, hay, n, pos, &s, &e);
    if (k <= 0) break;
    if (s == e && e == last_end) { /* handle_overlapping_empty_match */
      pos++;
      if (pos > n) break;
      k = regex_find(&f, &r, hay, n, pos, &s, &e);
      if (k <= 0) break;
    }
    if ((size_t)cnt < cap) { spans[2 * cnt] = (uint32_t)s; spans[2 * cnt + 1] = (uint32_t)e; }
    cnt++;
    pos = e; last_end = e;
  }
  return cnt;
}

static int contains(const uint8_t *h, size_t n, const uint8_t *needle, size_t m) {
  if (m == 0) return 1;
  for (size_t i = 0; i + m <= n; i++)
    if (memcmp(h + i, needle, m) == 0) return 1;
  return 0;
}
/* core/src/regex.rs:15-53.  returns 1 verified, 0 not verified, -1 bad DFA bytes */
static int process_regex_parts(const zo_regex_part *parts, size_t n_parts, const uint8_t *input,
                               size_t n, zo_result *out) {
  for (size_t pi = 0; pi < n_parts; pi++) {
    uint32_t span[2] = {0, 0};
    long cnt = zo_dfa_find_iter(parts[pi].fwd, parts[pi].fwd_len, parts[pi].bwd, parts[pi].bwd_len,
                                input, n, span, 1);
    uint32_t slot = out->n_parts;
    if (slot < ZO_MAX_PARTS) {
      out->parts[slot].match_count = cnt < 0 ? 0 : (uint32_t)cnt;
      out->parts[slot].start = span[0]; out->parts[slot].end = span[1];
      out->parts[slot].captures_ok = 0;
      out->n_parts++;
    }
    if (cnt < 0) return -1;
    if (cnt != 1) return 0;
    int ok = 1;
    if (parts[pi].captures) {
      size_t ml = span[1] - span[0];
      uint8_t *lossy = (uint8_t *)malloc(3 * ml + 1);
      size_t ll = zo_utf8_lossy(input + span[0], ml, lossy);
      for (size_t c = 0; c < parts[pi].n_captures; c++)
        if (!contains(lossy, ll, (const uint8_t *)parts[pi].captures[c], strlen(parts[pi].captures[c]))) { ok = 0; break; }
      free(lossy);
    }
    if (slot < ZO_MAX_PARTS) out->parts[slot].captures_ok = (uint32_t)ok;
    if (!ok) return 0;
  }
  return 1;
}

/* ------------------------------------------------------------------ core/src/circuits.rs */
void zo_verify_email(const zo_email *e, int64_t now_unix, zo_result *out) {
  memset(out, 0, sizeof *out);
  out->dkim_detail = ZO_DKIM_NEUTRAL;
  mail_t m;
  if (mail_parse(e->raw_email, e->raw_len, &m)) { out->status = ZO_ERR_MAIL_PARSE; return; }
  /* DkimPublicKey::try_from_bytes */
  if (strcmp(e->key_type, "rsa") == 0) {
    if (zo_parse_rsa_der(e->key, e->key_len, NULL, NULL, NULL)) { out->status = ZO_ERR_KEY; free(m.q); return; }
  } else if (strcmp(e->key_type, "ed25519") == 0) {
    out->status = e->key_len == 32 ? ZO_ERR_UNSUPPORTED : ZO_ERR_KEY; free(m.q); return;
  } else { out->status = ZO_ERR_KEY; free(m.q); return; }
  verify_dkim(e, &m, e->key, e->key_len, now_unix, out);
  free(m.q);
  if (out->status == ZO_ERR_UNSUPPORTED) return;
  if (out->dkim_detail != ZO_DKIM_PASS) { out->status = ZO_ERR_DKIM_FAIL; return; }
  zo_sha256((const uint8_t *)e->from_domain, e->from_domain_len, out->from_domain_hash);
  zo_sha256(e->key, e->key_len, out->public_key_hash);
}

void zo_verify_email_with_regex(const zo_email *e, const zo_regex_part *header_parts,
                                size_t n_header, const zo_regex_part *body_parts, size_t n_body,
                                int64_t now_unix, zo_result *out) {
  zo_verify_email(e, now_unix, out);
  if (out->status != ZO_OK) return;
  uint8_t *hdr = NULL, *body = NULL;
  size_t hl = 0, bl = 0;
  if (zo_canonicalize_signed_email(e->raw_email, e->raw_len, now_unix, &hdr, &hl, &body, &bl)) {
    out->status = ZO_ERR_CANONICALIZE;
    return;
  }
  uint8_t *clean = (uint8_t *)malloc(bl + 1);
  zo_qp_clean(body, bl, clean);
  if (header_parts) {
    int r = process_regex_parts(header_parts, n_header, hdr, hl, out);
    if (r < 0) out->status = ZO_ERR_BAD_DFA;
    else if (r == 0) out->status = ZO_ERR_REGEX_HEADER;
  }
  if (out->status == ZO_OK && body_parts) {
    int r = process_regex_parts(body_parts, n_body, clean, bl, out);
    if (r < 0) out->status = ZO_ERR_BAD_DFA;
    else if (r == 0) out->status = ZO_ERR_REGEX_BODY;
  }
  free(clean); free(hdr); free(body);
}

/* ------------------------------------------------------------------ multi-thread batch driver */
typedef struct {
  const zo_email *emails; size_t n;
  const zo_regex_part *hp; size_t nh; const zo_regex_part *bp; size_t nb;
  int64_t now; int use_openssl; zo_result *out;
  size_t *next; pthread_mutex_t *mu;
} mt_arg;
static void *mt_worker(void *p) {
  mt_arg *a = (mt_arg *)p;
  g_use_openssl = a->use_openssl;
  for (;;) {
    pthread_mutex_lock(a->mu);
    size_t lo = *a->next;
    size_t hi = lo + 64 < a->n ? lo + 64 : a->n;
    *a->next = hi;
    pthread_mutex_unlock(a->mu);
    if (lo >= a->n) break;
    for (size_t i = lo; i < hi; i++) {
      if (a->hp || a->bp) zo_verify_email_with_regex(&a->emails[i], a->hp, a->nh, a->bp, a->nb, a->now, &a->out[i]);
      else zo_verify_email(&a->emails[i], a->now, &a->out[i]);
    }
  }
  g_use_openssl = 0;
  return NULL;
}
int zo_verify_batch_mt(const zo_email *emails, size_t n, const zo_regex_part *header_parts,
                       size_t n_header, const zo_regex_part *body_parts, size_t n_body,
                       int64_t now_unix, int n_threads, int use_openssl, zo_result *out) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 512) n_threads = 512;
#ifndef ZO_WITH_OPENSSL
  if (use_openssl) return -1;
#endif
  size_t next = 0;
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  mt_arg a = {emails, n, header_parts, n_header, body_parts, n_body, now_unix, use_openssl, out, &next, &mu};
  pthread_t th[512];
  for (int i = 1; i < n_threads; i++) pthread_create(&th[i], NULL, mt_worker, &a);
  mt_worker(&a);
  for (int i = 1; i < n_threads; i++) pthread_join(th[i], NULL);
  return 0;
}
int zo_has_openssl(void) {
#ifdef ZO_WITH_OPENSSL
  return 1;
#else
  return 0;
#endif
}
