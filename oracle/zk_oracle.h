/*
 * zk_oracle.h — CPU ORACLE for the zkemail_core::verify_email / verify_email_with_regex hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (zkemail.rs_b200/, libzkemail_b200.so) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" with respect to the Rust binary.  The reference
 * (/root/reference, zkemail/zkemail.rs) cannot be compiled here (no Rust toolchain) and its
 * arithmetic lives in crates that are not vendored (cfdkim 0.3.3 @75af99fb, rsa 0.9.6,
 * sha2 0.10.9, regex-automata 0.4.9, mailparse 0.15.0, base64 0.21.7 — see Cargo.lock).  The
 * reference holds no golden vectors for this path (its only test is a network fetch,
 * helpers/src/dkim.rs:118-146).  This restatement is therefore pinned by (i) public KATs
 * (FIPS 180-4, RFC 6376 §3.4.5, RFC 8017 DigestInfo), (ii) differential tests against hashlib,
 * `cryptography`, Python big-int pow and an independent Python canonicaliser (tests/), and
 * (iii) the committed fixtures under tests/golden/.
 *
 * Every function cites the reference file:line (relative to /root/reference) or the dependency
 * behaviour it follows.
 */
#ifndef ZK_ORACLE_H
#define ZK_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes: one per panic site of the reference (SURVEY.md §8b "Error convention") ---- */
enum {
  ZO_OK = 0,
  ZO_ERR_MAIL_PARSE = 1,     /* core/src/email.rs:26   parse_mail(..).unwrap()                 */
  ZO_ERR_KEY = 2,            /* core/src/email.rs:28-29 DkimPublicKey::try_from_bytes().unwrap() */
  ZO_ERR_DKIM_FAIL = 3,      /* core/src/circuits.rs:13 assert!(verified)                       */
  ZO_ERR_NULL_EXTERNAL = 4,  /* core/src/circuits.rs:24 expect("Value cannot be null")          */
  ZO_ERR_CANONICALIZE = 5,   /* core/src/circuits.rs:35 canonicalize_signed_email().unwrap()    */
  ZO_ERR_REGEX_HEADER = 6,   /* core/src/circuits.rs:45 assert!(verified) header parts          */
  ZO_ERR_REGEX_BODY = 7,     /* core/src/circuits.rs:54 assert!(verified) body parts            */
  ZO_ERR_BAD_DFA = 8,        /* core/src/regex.rs:32-33 DFA::from_bytes().unwrap()              */
  ZO_ERR_UNSUPPORTED = 9     /* ed25519 key / rsa-sha1: the engine declines (never mis-verifies) */
};

/* ---- DKIMResult detail (cfdkim DKIMError kinds; only "pass" is observable in the reference) ---- */
enum {
  ZO_DKIM_PASS = 0,
  ZO_DKIM_NEUTRAL = 1,
  ZO_DKIM_SYNTAX = 2,
  ZO_DKIM_MISSING_TAG = 3,
  ZO_DKIM_VERSION = 4,
  ZO_DKIM_DOMAIN_MISMATCH = 5,
  ZO_DKIM_FROM_NOT_SIGNED = 6,
  ZO_DKIM_QUERY_METHOD = 7,
  ZO_DKIM_EXPIRED = 8,
  ZO_DKIM_CANON_TYPE = 9,
  ZO_DKIM_HASH_ALGO = 10,
  ZO_DKIM_BODY_HASH = 11,
  ZO_DKIM_SIG_SYNTAX = 12,
  ZO_DKIM_SIG_MISMATCH = 13,
  ZO_DKIM_LENGTH_TAG = 14,
  ZO_DKIM_ALGO_KEY_MISMATCH = 15
};

#define ZO_MAX_PARTS 16

typedef struct {
  int32_t status;          /* ZO_OK or first panic site in program order */
  int32_t dkim_detail;     /* ZO_DKIM_* */
  uint8_t body_hash[32];   /* SHA-256 of canonical body of the deciding signature (zeros if none) */
  uint8_t header_hash[32]; /* SHA-256 of the header preimage of the deciding signature */
  uint8_t from_domain_hash[32]; /* core/src/circuits.rs:16 */
  uint8_t public_key_hash[32];  /* core/src/circuits.rs:17 */
  uint8_t bh_ok, rsa_ok;
  uint8_t pad[2];
  uint32_t n_parts;        /* header parts then body parts */
  struct {
    uint32_t match_count, start, end; /* first match span (valid if match_count>=1) */
    uint32_t captures_ok;             /* all captures are substrings of the match */
  } parts[ZO_MAX_PARTS];
} zo_result;

typedef struct {
  const uint8_t *fwd; size_t fwd_len;   /* ZDF1 tables (see include/zkemail_b200.h) */
  const uint8_t *bwd; size_t bwd_len;
  const char *const *captures; size_t n_captures; /* NULL => captures: None */
} zo_regex_part;

typedef struct {
  const char *from_domain; size_t from_domain_len;
  const uint8_t *raw_email; size_t raw_len;
  const uint8_t *key; size_t key_len;
  const char *key_type; /* NUL-terminated */
} zo_email;

/* primitives (exposed so tests can pin each one) */
void zo_sha256(const uint8_t *data, size_t len, uint8_t out[32]);
size_t zo_base64_encode(const uint8_t *in, size_t n, char *out);          /* returns chars written */
long zo_base64_decode(const uint8_t *in, size_t n, uint8_t *out);         /* -1 on error */
/* out = base^exp mod mod; big-endian byte strings, out has mod_len bytes. returns 0 ok */
int zo_modexp(const uint8_t *base, size_t blen, uint64_t exp, const uint8_t *mod, size_t mlen,
              uint8_t *out);
/* PKCS#1 RSAPublicKey DER -> n (big endian, minimal) and e. returns 0 ok, <0 on reject */
int zo_parse_rsa_der(const uint8_t *der, size_t len, uint8_t *n_out, size_t *n_len, uint64_t *e);
/* rsa 0.9.6 Pkcs1v15Sign::<Sha256>::verify. returns 1 pass, 0 fail, <0 key error */
int zo_rsa_verify_sha256(const uint8_t *der, size_t der_len, const uint8_t hash[32],
                         const uint8_t *sig, size_t sig_len);

/* canonicalisation pieces (cfdkim canonicalization.rs) */
size_t zo_canon_body_relaxed(const uint8_t *in, size_t n, uint8_t *out); /* out cap >= n+2 */
size_t zo_canon_body_simple(const uint8_t *in, size_t n, uint8_t *out);  /* out cap >= n+2 */
size_t zo_canon_header_relaxed(const uint8_t *key, size_t klen, const uint8_t *val, size_t vlen,
                               uint8_t *out); /* out cap >= 2*klen+vlen+3 */
size_t zo_canon_header_simple(const uint8_t *key, size_t klen, const uint8_t *val, size_t vlen,
                              uint8_t *out);  /* out cap >= 2*klen+vlen+4 */
/* mailparse::parse_headers: fills up to cap (key_off,key_len,val_off,val_len) quads.
 * returns header count or -1 on the parse errors that make parse_mail fail.  *body_off = index
 * just past the blank line. */
long zo_parse_headers(const uint8_t *raw, size_t n, uint32_t *quads, size_t cap, size_t *body_off);
/* String::from_utf8_lossy; out cap >= 3*n */
size_t zo_utf8_lossy(const uint8_t *in, size_t n, uint8_t *out);
/* core/src/email.rs:61-86; out has n bytes (zero padded). returns cleaned length before padding */
size_t zo_qp_clean(const uint8_t *in, size_t n, uint8_t *out);

/* cfdkim::canonicalize_signed_email (core/src/circuits.rs:34-35).  Buffers malloc'ed, caller frees
 * with zo_free. returns 0 ok, nonzero => the reference's unwrap() panics. */
int zo_canonicalize_signed_email(const uint8_t *raw, size_t n, int64_t now_unix, uint8_t **hdr,
                                 size_t *hdr_len, uint8_t **body, size_t *body_len);
void zo_free(void *p);

/* regex-automata dfa::regex::Regex::find_iter over ZDF1 tables (core/src/regex.rs:32-36).
 * Writes up to cap (start,end) pairs; returns total match count, or -1 for bad tables. */
long zo_dfa_find_iter(const uint8_t *fwd, size_t fwd_len, const uint8_t *bwd, size_t bwd_len,
                      const uint8_t *hay, size_t n, uint32_t *spans, size_t cap);

/* core/src/circuits.rs:9-29 (without external_inputs flattening, which is host string glue) */
void zo_verify_email(const zo_email *e, int64_t now_unix, zo_result *out);
/* core/src/circuits.rs:31-68 */
void zo_verify_email_with_regex(const zo_email *e, const zo_regex_part *header_parts,
                                size_t n_header, const zo_regex_part *body_parts, size_t n_body,
                                int64_t now_unix, zo_result *out);

/* batch driver over host threads (CPU baseline leg of bench.py; rayon-over-all-cores stand-in).
 * use_openssl!=0 routes SHA-256 and the RSA public op through libcrypto (if built with it). */
int zo_verify_batch_mt(const zo_email *emails, size_t n, const zo_regex_part *header_parts,
                       size_t n_header, const zo_regex_part *body_parts, size_t n_body,
                       int64_t now_unix, int n_threads, int use_openssl, zo_result *out);
int zo_has_openssl(void);

#ifdef __cplusplus
}
#endif
#endif
