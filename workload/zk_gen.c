/*
 * zk_gen.c — TEST/BENCH INFRASTRUCTURE: fast synthetic DKIM-signed mail generator (C + OpenSSL,
 * multi-threaded).  Offline stand-in for helpers/src/generator.rs:11-53 (which needs DNS/HTTP for
 * the key): produces RFC 5322 messages with one relaxed/relaxed rsa-sha256 DKIM-Signature, signed
 * with freshly generated RSA keys (SURVEY.md §8d recipe C2-C5).  It is a THIRD, independent signer
 * (after zkemail.rs_b200/synth.py and the oracle): it writes the relaxed header preimage directly
 * from its own header template, so a canonicalisation bug in the oracle or the engine shows up as
 * a signature that does not verify.  Never linked into the product.
 */
#include <openssl/bn.h>
#include <openssl/evp.h>
#include <openssl/rsa.h>
#include <openssl/sha.h>
#include <openssl/x509.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ZG_MAX_KEYS 4096
#define ZG_DER_STRIDE 512
#define ZG_DOM_STRIDE 64

typedef struct {
  EVP_PKEY *pkey[ZG_MAX_KEYS];
  int bits[ZG_MAX_KEYS];
  int n_keys;
} zg_keys;

static uint64_t splitmix(uint64_t *s) {
  uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

typedef struct { zg_keys *k; int lo, hi; } kg_arg;
static void *kg_worker(void *p) {
  kg_arg *a = (kg_arg *)p;
  for (int i = a->lo; i < a->hi; i++) {
    EVP_PKEY_CTX *ctx = EVP_PKEY_CTX_new_id(EVP_PKEY_RSA, NULL);
    EVP_PKEY_keygen_init(ctx);
    EVP_PKEY_CTX_set_rsa_keygen_bits(ctx, a->k->bits[i]);
    EVP_PKEY *pk = NULL;
    EVP_PKEY_keygen(ctx, &pk);
    EVP_PKEY_CTX_free(ctx);
    a->k->pkey[i] = pk;
  }
  return NULL;
}

/* Generates n2048 + n1024 fresh keys (e = 65537).  der_out: n x ZG_DER_STRIDE bytes (PKCS#1
 * RSAPublicKey DER, the PublicKey.key contract of helpers/src/dkim.rs:50), der_len: n. */
void *zg_keys_create(int n2048, int n1024, int n_threads, uint8_t *der_out, uint32_t *der_len) {
  int n = n2048 + n1024;
  if (n <= 0 || n > ZG_MAX_KEYS) return NULL;
  zg_keys *k = (zg_keys *)calloc(1, sizeof *k);
  k->n_keys = n;
  for (int i = 0; i < n; i++) k->bits[i] = i < n2048 ? 2048 : 1024;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 64) n_threads = 64;
  pthread_t th[64];
  kg_arg args[64];
  for (int t = 0; t < n_threads; t++) {
    args[t].k = k; args[t].lo = (int)((long)n * t / n_threads); args[t].hi = (int)((long)n * (t + 1) / n_threads);
    pthread_create(&th[t], NULL, kg_worker, &args[t]);
  }
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  for (int i = 0; i < n; i++) {
    if (!k->pkey[i]) return NULL;
    BIGNUM *bn_n = NULL, *bn_e = NULL;
    EVP_PKEY_get_bn_param(k->pkey[i], "n", &bn_n);
    EVP_PKEY_get_bn_param(k->pkey[i], "e", &bn_e);
    /* SEQUENCE { INTEGER n, INTEGER e } written by hand (minimal DER) */
    uint8_t nb[520], eb[16];
    int nl = BN_bn2bin(bn_n, nb + 1), el = BN_bn2bin(bn_e, eb + 1);
    uint8_t *np = nb + 1, *ep = eb + 1;
    if (np[0] & 0x80) { np--; np[0] = 0; nl++; }
    if (ep[0] & 0x80) { ep--; ep[0] = 0; el++; }
    uint8_t body[600];
    int o = 0;
    body[o++] = 0x02;
    if (nl < 128) body[o++] = (uint8_t)nl;
    else if (nl < 256) { body[o++] = 0x81; body[o++] = (uint8_t)nl; }
    else { body[o++] = 0x82; body[o++] = (uint8_t)(nl >> 8); body[o++] = (uint8_t)nl; }
    memcpy(body + o, np, nl); o += nl;
    body[o++] = 0x02; body[o++] = (uint8_t)el;
    memcpy(body + o, ep, el); o += el;
    uint8_t *d = der_out + (size_t)i * ZG_DER_STRIDE;
    int h = 0;
    d[h++] = 0x30;
    if (o < 128) d[h++] = (uint8_t)o;
    else if (o < 256) { d[h++] = 0x81; d[h++] = (uint8_t)o; }
    else { d[h++] = 0x82; d[h++] = (uint8_t)(o >> 8); d[h++] = (uint8_t)o; }
    memcpy(d + h, body, o);
    der_len[i] = (uint32_t)(h + o);
    BN_free(bn_n); BN_free(bn_e);
  }
  return k;
}
void zg_keys_destroy(void *kp) {
  zg_keys *k = (zg_keys *)kp;
  if (!k) return;
  for (int i = 0; i < k->n_keys; i++) EVP_PKEY_free(k->pkey[i]);
  free(k);
}

static const char B64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
static size_t b64enc(const uint8_t *in, size_t n, char *out) {
  size_t o = 0, i = 0;
  for (; i + 3 <= n; i += 3) {
    uint32_t v = ((uint32_t)in[i] << 16) | ((uint32_t)in[i + 1] << 8) | in[i + 2];
    out[o++] = B64[v >> 18]; out[o++] = B64[(v >> 12) & 63]; out[o++] = B64[(v >> 6) & 63]; out[o++] = B64[v & 63];
  }
  if (n - i == 1) { uint32_t v = (uint32_t)in[i] << 16; out[o++] = B64[v >> 18]; out[o++] = B64[(v >> 12) & 63]; out[o++] = '='; out[o++] = '='; }
  else if (n - i == 2) { uint32_t v = ((uint32_t)in[i] << 16) | ((uint32_t)in[i + 1] << 8); out[o++] = B64[v >> 18]; out[o++] = B64[(v >> 12) & 63]; out[o++] = B64[(v >> 6) & 63]; out[o++] = '='; }
  return o;
}

/* body of exactly `len` bytes: printable lines of 40..76 chars + CRLF, canonical-stable under
 * relaxed canonicalisation (no TAB, no double SP, no SP before CRLF, ends with one CRLF).
 * token (optional) is planted as its own line; qp != 0 inserts "=\r\n" soft breaks (one inside the
 * token, some inside other lines). */
static void gen_body(uint64_t *rs, uint8_t *out, size_t len, const char *token, int qp) {
  static const char AL[] = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789 ,.;:-_()";
  const size_t NA = sizeof(AL) - 1;
  if (len < 3) { memset(out, 'x', len); return; }
  char tok[160];
  size_t tl = 0;
  if (token) {
    size_t L = strlen(token);
    if (qp && L > 6) {
      size_t cut = 1 + splitmix(rs) % (L - 2);
      memcpy(tok, token, cut); memcpy(tok + cut, "=\r\n", 3); memcpy(tok + cut + 3, token + cut, L - cut);
      tl = L + 3;
    } else { memcpy(tok, token, L); tl = L; }
    tok[tl++] = '\r'; tok[tl++] = '\n';
    if (len < tl + 3) { token = NULL; tl = 0; }
  }
  size_t tok_at = token ? splitmix(rs) % (len - tl - 2) : (size_t)-1;
  size_t o = 0;
  while (o < len) {
    size_t rem = len - o;
    if (token && o >= tok_at && rem >= tl && (rem - tl == 0 || rem - tl >= 3)) {
      memcpy(out + o, tok, tl); o += tl; token = NULL;
      continue;
    }
    size_t budget = rem - (token ? tl : 0);
    if (budget < 3) { budget = rem; }
    size_t ll = 40 + splitmix(rs) % 37;
    if (ll > budget - 2) ll = budget - 2;
    size_t left = budget - (ll + 2);
    if (left == 1 || left == 2) { if (ll + left <= 76) ll += left; else ll -= (3 - left); }
    if (ll < 1) ll = 1;
    int soft = qp && ll > 16 && (splitmix(rs) % 10) < 3;
    uint8_t prev = 0;
    for (size_t i = 0; i < ll; i++) {
      uint8_t c = (uint8_t)AL[splitmix(rs) % NA];
      if (c == ' ' && (i == 0 || i + 1 == ll || prev == ' ')) c = 'x';
      if (soft && i + 3 < ll && i == ll / 2) { out[o + i] = '='; out[o + i + 1] = '\r'; out[o + i + 2] = '\n'; i += 2; prev = '\n'; continue; }
      out[o + i] = c; prev = c;
    }
    o += ll;
    out[o++] = '\r'; out[o++] = '\n';
  }
}

typedef struct {
  zg_keys *keys;
  uint64_t seed;
  size_t n, lo, hi;
  const uint32_t *body_len; const uint32_t *key_idx; const uint8_t *neg_kind; const uint64_t *raw_off;
  uint8_t *raw; uint32_t *raw_len;
  int token, qp_percent;
  int rc;
} gen_arg;

static void *gen_worker(void *p) {
  gen_arg *a = (gen_arg *)p;
  EVP_MD_CTX *md = EVP_MD_CTX_new();
  char *pre = (char *)malloc(4096), *hdr = (char *)malloc(8192);
  for (size_t i = a->lo; i < a->hi; i++) {
    uint64_t rs = a->seed ^ (0xD1C1ull * (i + 1));
    splitmix(&rs);
    const int ki = (int)a->key_idx[i];
    const size_t bl = a->body_len[i];
    uint8_t *raw = a->raw + a->raw_off[i];
    char user[9], subj[25], dom[48], tokbuf[64];
    for (int k = 0; k < 8; k++) user[k] = (char)('a' + splitmix(&rs) % 26);
    user[8] = 0;
    for (int k = 0; k < 24; k++) subj[k] = (char)('a' + splitmix(&rs) % 26);
    subj[24] = 0;
    snprintf(dom, sizeof dom, "d%d.example.com", ki);
    const char *token = NULL;
    if (a->token) {
      static const char TA[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789";
      char id[13];
      for (int k = 0; k < 12; k++) id[k] = TA[splitmix(&rs) % 36];
      id[12] = 0;
      snprintf(tokbuf, sizeof tokbuf, "Transaction ID: %s", id);
      token = tokbuf;
    }
    const int qp = a->qp_percent > 0 && (int)(splitmix(&rs) % 100) < a->qp_percent;
    /* headers as they appear on the wire */
    int hl = snprintf(hdr, 8192,
                      "Received: from mx.%s by relay.example.net; Mon, 1 Jan 2024 00:00:00 +0000\r\n"
                      "From: %c%s <%s@%s>\r\n"
                      "To: recipient%zu@example.org\r\n"
                      "Subject: Order %s update %zu\r\n"
                      "Date: Mon, 01 Jan 2024 00:00:00 +0000\r\n"
                      "Message-ID: <%08zu.%s@%s>\r\n"
                      "MIME-Version: 1.0\r\n"
                      "Content-Type: text/plain; charset=us-ascii\r\n",
                      dom, user[0] - 32, user + 1, user, dom, i, subj, i, i, user, dom);
    /* body + bh */
    size_t sig_b64_len = (size_t)(a->keys->bits[ki] / 8 + 2) / 3 * 4;
    /* DKIM-Signature header: "DKIM-Signature: v=1; ...;\r\n\th=...;\r\n\tbh=...;\r\n\tb=" + folded sig */
    uint8_t *body_tmp = (uint8_t *)malloc(bl + 8);
    gen_body(&rs, body_tmp, bl, token, qp);
    uint8_t bh[32];
    SHA256(body_tmp, bl, bh);
    char bh64[48];
    size_t bhl = b64enc(bh, 32, bh64);
    bh64[bhl] = 0;
    /* relaxed header preimage (RFC 6376 3.4.2 applied to the template above) */
    int pl = snprintf(pre, 4096,
                      "from:%c%s <%s@%s>\r\n"
                      "to:recipient%zu@example.org\r\n"
                      "subject:Order %s update %zu\r\n"
                      "date:Mon, 01 Jan 2024 00:00:00 +0000\r\n"
                      "message-id:<%08zu.%s@%s>\r\n"
                      "dkim-signature:v=1; a=rsa-sha256; c=relaxed/relaxed; d=%s; s=sel1; h=from:to:subject:date:message-id; bh=%s; b=",
                      user[0] - 32, user + 1, user, dom, i, subj, i, i, user, dom, dom, bh64);
    uint8_t sig[512];
    size_t sl = sizeof sig;
    EVP_MD_CTX_reset(md);
    if (EVP_DigestSignInit(md, NULL, EVP_sha256(), NULL, a->keys->pkey[ki]) != 1 ||
        EVP_DigestSign(md, sig, &sl, (const uint8_t *)pre, (size_t)pl) != 1) { a->rc = 1; free(body_tmp); break; }
    char s64[700];
    size_t s64l = b64enc(sig, sl, s64);
    (void)sig_b64_len;
    const uint8_t neg = a->neg_kind[i];
    if (neg == 2) s64[7] = s64[7] == 'A' ? 'B' : 'A'; /* sig flip */
    size_t o = 0;
    o += (size_t)snprintf((char *)raw + o, 1024,
                          "DKIM-Signature: v=1; a=rsa-sha256; c=relaxed/relaxed; d=%s; s=sel1;\r\n"
                          "\th=from:to:subject:date:message-id;\r\n\tbh=%s;\r\n\tb=", dom, bh64);
    size_t first = 60;
    for (size_t q = 0; q < s64l;) {
      size_t w = q == 0 ? first : 72;
      if (w > s64l - q) w = s64l - q;
      if (q) { raw[o++] = '\r'; raw[o++] = '\n'; raw[o++] = '\t'; }
      memcpy(raw + o, s64 + q, w); o += w; q += w;
    }
    raw[o++] = '\r'; raw[o++] = '\n';
    memcpy(raw + o, hdr, (size_t)hl); o += (size_t)hl;
    raw[o++] = '\r'; raw[o++] = '\n';
    memcpy(raw + o, body_tmp, bl);
    if (neg == 1 && bl > 0) { /* body flip: change one letter into another letter */
      size_t at = splitmix(&rs) % bl;
      for (size_t t = 0; t < bl; t++) {
        uint8_t c = raw[o + (at + t) % bl];
        if ((c >= 'a' && c <= 'y') || (c >= 'A' && c <= 'Y')) { raw[o + (at + t) % bl] = c + 1; break; }
      }
    }
    o += bl;
    a->raw_len[i] = (uint32_t)o;
    free(body_tmp);
  }
  EVP_MD_CTX_free(md);
  free(pre); free(hdr);
  return NULL;
}

/* Upper bound of the raw size of an email with the given body length. */
size_t zg_raw_bound(size_t body_len) { return body_len + 1400; }

/* Fills n emails.  body_len[i]: canonical body bytes; key_idx[i]: signing key (the from_domain is
 * "d<key_idx>.example.com"); neg_kind[i]: 0 valid, 1 body flip, 2 signature flip (3 = "wrong key"
 * is applied by the caller by handing the verifier another key); raw_off[i]: where to write in
 * `raw` (>= zg_raw_bound(body_len[i]) bytes available); raw_len[i] out. */
int zg_generate(void *kp, uint64_t seed, size_t n, const uint32_t *body_len, const uint32_t *key_idx,
                const uint8_t *neg_kind, const uint64_t *raw_off, uint8_t *raw, uint32_t *raw_len, int token,
                int qp_percent, int n_threads) {
  zg_keys *k = (zg_keys *)kp;
  if (!k) return 1;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256];
  gen_arg args[256];
  for (int t = 0; t < n_threads; t++) {
    gen_arg g = {k, seed, n, n * (size_t)t / n_threads, n * (size_t)(t + 1) / n_threads, body_len, key_idx, neg_kind, raw_off, raw, raw_len, token, qp_percent, 0};
    args[t] = g;
    pthread_create(&th[t], NULL, gen_worker, &args[t]);
  }
  int rc = 0;
  for (int t = 0; t < n_threads; t++) { pthread_join(th[t], NULL); rc |= args[t].rc; }
  return rc;
}

/* Packs the generated messages tightly (each start `align`-byte aligned) by moving them towards the front of
 * `raw`; rewrites raw_off.  Returns the number of bytes in use. */
size_t zg_compact(uint8_t *raw, uint64_t *raw_off, const uint32_t *raw_len, size_t n, size_t align) {
  size_t o = 0;
  for (size_t i = 0; i < n; i++) {
    o = (o + align - 1) / align * align;
    if (raw_off[i] != o) memmove(raw + o, raw + raw_off[i], raw_len[i]);
    raw_off[i] = o;
    o += raw_len[i];
  }
  return o;
}
