/*
 * zkemail_b200.h — C ABI of libzkemail_b200.so: the B200-native batched replacement for the
 * native execution of zkemail_core::verify_email / verify_email_with_regex.
 *
 * The reference has no FFI of its own: callers link the Rust crate and call
 *     verify_email(&Email) -> EmailVerifierOutput                       core/src/circuits.rs:9
 *     verify_email_with_regex(&EmailWithRegex) -> EmailWith..Output     core/src/circuits.rs:31
 * so the entry points below are what a `zkemail-b200-sys` crate would bind (INTEGRATION.md shows
 * the Rust side).  Plain pointers and sizes only; caller owns every input and output buffer; no
 * function throws or aborts across the ABI: each returns a ZKB_E_* code (0 = ok), and every
 * panic site of the reference becomes a per-email `status` (never a process abort).
 *
 * There is NO CPU fallback: every hash, modular exponentiation and DFA scan runs in sm_100a
 * kernels.  If no CUDA device is usable, zkb_engine_create fails with ZKB_E_NO_DEVICE.
 */
#ifndef ZKEMAIL_B200_H
#define ZKEMAIL_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ZKB_ABI_VERSION 1

/* ---- library error codes (return values) ---- */
enum {
  ZKB_OK = 0,
  ZKB_E_INVALID = 1,   /* bad argument */
  ZKB_E_NO_DEVICE = 2, /* no usable CUDA device (there is no CPU fallback) */
  ZKB_E_CUDA = 3,      /* CUDA runtime error (message on stderr) */
  ZKB_E_NOMEM = 4,
  ZKB_E_REGEX = 5,     /* pattern rejected by the regex compiler / bad table bytes */
  ZKB_E_UNSUPPORTED = 6
};

/* ---- per-email status: one value per panic site of the reference, in program order ---- */
enum {
  ZKB_ST_OK = 0,
  ZKB_ST_MAIL_PARSE = 1,    /* core/src/email.rs:26    parse_mail(..).unwrap()                   */
  ZKB_ST_KEY = 2,           /* core/src/email.rs:28-29 DkimPublicKey::try_from_bytes(..).unwrap() */
  ZKB_ST_DKIM_FAIL = 3,     /* core/src/circuits.rs:13 assert!(verified)                         */
  ZKB_ST_NULL_EXTERNAL = 4, /* core/src/circuits.rs:24 expect("Value cannot be null") (wrappers)  */
  ZKB_ST_CANONICALIZE = 5,  /* core/src/circuits.rs:35 canonicalize_signed_email(..).unwrap()     */
  ZKB_ST_REGEX_HEADER = 6,  /* core/src/circuits.rs:45 assert!(verified)                         */
  ZKB_ST_REGEX_BODY = 7,    /* core/src/circuits.rs:54 assert!(verified)                         */
  ZKB_ST_BAD_DFA = 8,       /* core/src/regex.rs:32-33 DFA::from_bytes(..).unwrap()               */
  ZKB_ST_UNSUPPORTED = 9    /* ed25519 key or rsa-sha1 signature: declined, never mis-verified   */
};

/* ---- DKIMResult detail (cfdkim DKIMError kind of the deciding signature) ---- */
enum {
  ZKB_DKIM_PASS = 0, ZKB_DKIM_NEUTRAL = 1, ZKB_DKIM_SYNTAX = 2, ZKB_DKIM_MISSING_TAG = 3,
  ZKB_DKIM_VERSION = 4, ZKB_DKIM_DOMAIN_MISMATCH = 5, ZKB_DKIM_FROM_NOT_SIGNED = 6,
  ZKB_DKIM_QUERY_METHOD = 7, ZKB_DKIM_EXPIRED = 8, ZKB_DKIM_CANON_TYPE = 9,
  ZKB_DKIM_HASH_ALGO = 10, ZKB_DKIM_BODY_HASH = 11, ZKB_DKIM_SIG_SYNTAX = 12,
  ZKB_DKIM_SIG_MISMATCH = 13, ZKB_DKIM_LENGTH_TAG = 14, ZKB_DKIM_ALGO_KEY_MISMATCH = 15
};

#define ZKB_MAX_PARTS 16

/* Borrowed view of one `Email` (core/src/structs.rs:49-54).  external_inputs are pure host string
 * glue (core/src/circuits.rs:18-27) and stay in the language wrapper. */
typedef struct {
  const char *from_domain;  size_t from_domain_len;
  const uint8_t *raw_email; size_t raw_email_len;
  const uint8_t *key;       size_t key_len;      /* PKCS#1 RSAPublicKey DER (PublicKey.key)   */
  const char *key_type;     size_t key_type_len; /* "rsa" | "ed25519"    (PublicKey.key_type) */
} zkb_email_view;

/* Per-email result record (fixed size POD; identical layout to the oracle's zo_result). */
typedef struct {
  int32_t status;               /* ZKB_ST_* */
  int32_t dkim_detail;          /* ZKB_DKIM_* */
  uint8_t body_hash[32];        /* SHA-256 of the canonical body of the deciding signature   */
  uint8_t header_hash[32];      /* SHA-256 of the signed-header preimage                     */
  uint8_t from_domain_hash[32]; /* EmailVerifierOutput.from_domain_hash (circuits.rs:16)     */
  uint8_t public_key_hash[32];  /* EmailVerifierOutput.public_key_hash  (circuits.rs:17)     */
  uint8_t bh_ok, rsa_ok;
  uint8_t pad[2];
  uint32_t n_parts;             /* header parts then body parts, as far as evaluated */
  struct {
    uint32_t match_count, start, end; /* find_iter count and the first match span */
    uint32_t captures_ok;             /* every expected capture is a substring of the match */
  } parts[ZKB_MAX_PARTS];
} zkb_result;

/* One `CompiledRegex.verify_re` (core/src/structs.rs:16-27): forward + reverse dense DFA tables in
 * the ZDF1 layout documented below.  Captures are per email and passed to zkb_verify_batch. */
typedef struct {
  const uint8_t *fwd; size_t fwd_len;
  const uint8_t *bwd; size_t bwd_len;
} zkb_dfa_view;

/* Expected capture strings of one email: `CompiledRegex.captures` flattened over the parts of the
 * regex set (header parts first).  part_has_captures[p]==0 means `captures: None` for part p. */
typedef struct {
  uint32_t part;       /* index into the regex set (header parts, then body parts) */
  const char *s; size_t len;
} zkb_capture;
typedef struct {
  const zkb_capture *caps; size_t n_caps;
} zkb_email_captures;

/* zkb_options.flags: path switches for diagnosis and A/B runs (every path gives identical results).  The library reads
 * no environment variables. */
enum {
  ZKB_OPT_NO_DIRECT = 1u,          /* ignore registered caller memory: raw messages are staged like pageable ones   */
  ZKB_OPT_NO_DEVICE_FRONTEND = 2u, /* header parsing / canonicalisation / base64 on the host threads for every mail */
  ZKB_OPT_NO_STAGED_FRONTEND = 4u, /* pageable callers: host front end instead of staging raw bytes for the device  */
  ZKB_OPT_NO_OVERLAP = 8u,         /* resident batches: one stream instead of the two-stream schedule               */
  ZKB_OPT_PROFILE = 16u,           /* per-call host / stream time breakdown on stderr                               */
  ZKB_OPT_NO_SQR = 32u,            /* RSA-2048, e = 65537, four lanes: the plain kernel (18 interleaved multiplications)
                                      instead of the one with the dedicated Montgomery squaring */
  ZKB_OPT_ALL = 63u
};

typedef struct {
  int32_t device;        /* CUDA device ordinal */
  int32_t host_threads;  /* 0 = all hardware threads */
  int64_t now_unix;      /* clock for the x= tag check; 0 = wall clock (cfdkim behaviour) */
  uint64_t chunk_emails; /* emails per pipeline chunk; 0 = default */
  uint32_t flags;        /* ZKB_OPT_* bits, 0 = defaults */
  uint32_t rsa_lanes;    /* lanes cooperating on one signature (0 = default) */
} zkb_options;

typedef struct zkb_engine zkb_engine;
typedef struct zkb_regex_set zkb_regex_set;
typedef struct zkb_batch zkb_batch; /* a prepared batch resident in device memory */

int zkb_abi_version(void);
const char *zkb_strerror(int code);

/* Engine = device context, streams, pinned staging, device arenas, key table. One per device. */
int zkb_engine_create(const zkb_options *opt, zkb_engine **out);
void zkb_engine_destroy(zkb_engine *e);
/* Replaces the engine's ZKB_OPT_* bits; takes effect with the next call on the engine. */
int zkb_engine_set_flags(zkb_engine *e, uint32_t flags);

/* Optional zero-copy input path.  Registers caller memory holding raw messages with the CUDA driver
 * (page-locks it; cudaHostRegister).  Batches whose zkb_email_view.raw_email pointers all lie inside
 * one registered range are DMA'd to the device as they are - no host copy - and their bodies are
 * canonicalised on the device (canon.cuh) instead of on the host threads.  Results are identical
 * either way.  Unregister before freeing the memory. */
int zkb_host_register(zkb_engine *e, const void *p, size_t len);
int zkb_host_unregister(zkb_engine *e, const void *p);

/* Registers the DFAs of a RegexInfo (core/src/structs.rs:32-35): n_header header parts followed
 * by n_body body parts.  header_present/body_present==0 model `None` (the part list is skipped,
 * core/src/circuits.rs:39-56). Tables are validated (bad bytes => ZKB_E_REGEX, the per-email
 * equivalent being ZKB_ST_BAD_DFA) and uploaded once. */
int zkb_regex_set_create(zkb_engine *e, const zkb_dfa_view *parts, size_t n_header, size_t n_body,
                         int header_present, int body_present, zkb_regex_set **out);
void zkb_regex_set_destroy(zkb_regex_set *s);

/* Batch entry point: verify_email (regex==NULL) or verify_email_with_regex over n emails.
 * `captures` may be NULL (no capture checks) or point to n entries.  Synchronous; thread-safe
 * for distinct engines, serialised per engine.  Replaces core/src/circuits.rs:9 and :31. */
int zkb_verify_batch(zkb_engine *e, const zkb_email_view *emails, size_t n,
                     const zkb_regex_set *regex, const zkb_email_captures *captures,
                     zkb_result *out);
/* Thin wrapper: batch of one (the reference's call shape). */
int zkb_verify_one(zkb_engine *e, const zkb_email_view *email, const zkb_regex_set *regex,
                   const zkb_email_captures *captures, zkb_result *out);

/* Split form used by bench.py to time the device-resident pass separately:
 *   prepare = host parse/canonicalise/pack + H2D;  run = kernels only (inputs resident in HBM);
 *   fetch = D2H + host post-processing into zkb_result records. */
int zkb_batch_prepare(zkb_engine *e, const zkb_email_view *emails, size_t n,
                      const zkb_regex_set *regex, const zkb_email_captures *captures,
                      zkb_batch **out);
/* Raw-resident form: the RAW messages are made resident (DMA of the registered span, or one staged copy of each
 * pageable message) and every zkb_batch_run* starts at the device front end: header parsing, preimages, base64, body
 * canonicalisation, hashing, RSA, DFA scans and the result records all inside the run.  Messages the device declines
 * are re-run through the host front end by zkb_batch_fetch.  This is the "from raw bytes in HBM" number of bench.py. */
int zkb_batch_prepare_raw(zkb_engine *e, const zkb_email_view *emails, size_t n,
                          const zkb_regex_set *regex, const zkb_email_captures *captures,
                          zkb_batch **out);
int zkb_batch_run(zkb_batch *b);                 /* enqueue + synchronise */
int zkb_batch_run_async(zkb_batch *b);           /* enqueue only on the engine stream */
int zkb_batch_fetch(zkb_batch *b, zkb_result *out);
void zkb_batch_destroy(zkb_batch *b);
/* Work actually resident for this batch (for the roofline arithmetic in bench.py). */
typedef struct {
  uint64_t n_emails, n_candidates, n_sha_messages, sha_blocks, sha_bytes;
  uint64_t rsa_items_1024, rsa_items_2048, rsa_items_other, rsa_macs;
  uint64_t dfa_items, dfa_bytes, arena_bytes, h2d_bytes, d2h_bytes;
  uint64_t kernel_launches;          /* kernels one zkb_batch_run launches */
} zkb_batch_stats;
int zkb_batch_get_stats(const zkb_batch *b, zkb_batch_stats *out);
/* CUDA-event time (ms) of the kernels of the last zkb_batch_run, per kernel family, measured on
 * the engine stream: [0]=sha256 [1]=rsa [2]=dfa [3]=bh/finalize [4]=whole run */
int zkb_batch_last_timing(const zkb_batch *b, float ms[5]);
/* The same with n <= 8 slots: [5]=device front end + body canonicalisation (raw-resident batches) [6]=result records */
int zkb_batch_last_timing_ex(const zkb_batch *b, float *ms, size_t n);
void *zkb_engine_stream(zkb_engine *e);          /* cudaStream_t of the engine */
/* Bytes the last zkb_verify_batch moved host->device / device->host, and how many of its messages the
 * device front end declined (re-run through the host front end). */
int zkb_engine_last_batch_bytes(const zkb_engine *e, uint64_t *h2d, uint64_t *d2h, uint64_t *host_front_end_emails);
/* Device-resident verdict words of chunk `chunk` of a prepared batch (one u32 per signature candidate:
 * bit0 = RSA ok, bit1 = bh= ok), for the multi-GPU all-gather of verdict bits over NCCL without a
 * host round trip.  *n_chunks (optional) receives the number of chunks.  Valid until the batch is
 * destroyed; written by zkb_batch_run*. */
int zkb_batch_device_flags(const zkb_batch *b, size_t chunk, void **flags, size_t *n_flags,
                           size_t *n_chunks);

/* ---- regex compiler (stands in for helpers/src/regex.rs:7-51, which needs regex-automata) ----
 * Compiles `pattern` (Rust-regex syntax subset, Unicode/UTF-8 mode as DFARegex::new) to forward
 * and reverse ZDF1 tables.  On success *fwd and *bwd are malloc'ed; free with zkb_free. */
int zkb_regex_compile(const char *pattern, size_t pattern_len, uint8_t **fwd, size_t *fwd_len,
                      uint8_t **bwd, size_t *bwd_len, char *err, size_t err_cap);
void zkb_free(void *p);
/* `DFA.fwd` / `DFA.bwd` may also hold regex-automata's own serialisation (dense::DFA::to_bytes_little_endian with the
 * leading padding stripped, helpers/src/regex.rs:7-14): zkb_regex_set_create and zkb_dfa_scan_batch detect its label
 * and convert it on load.  This entry point exposes the conversion (host-only): wire bytes -> malloc'ed ZDF1 table
 * (free with zkb_free).  ZKB_E_REGEX when the blob does not validate — the engine's equivalent of
 * dense::DFA::from_bytes(..).unwrap() panicking (core/src/regex.rs:32-33). */
int zkb_regex_automata_to_zdf(const uint8_t *wire, size_t wire_len, int reverse, uint8_t **zdf, size_t *zdf_len);

/* ---- multi-GPU ----------------------------------------------------------------------------
 * verify_email / verify_email_with_regex are pure per-email functions (core/src/circuits.rs:9,31): the batch shards by
 * email into contiguous, cost-balanced ranges, one engine per device, regex tables replicated.  The only exchange is an
 * NCCL all-gather of the fixed-size result records (the first 144 bytes of zkb_result + 16 bytes per regex part:
 * verdict, hashes, match spans), issued by this library on the engine streams.  NCCL is loaded on first use. */

/* Contiguous shards of roughly equal estimated cost: shard k = emails [bounds[k], bounds[k+1]); bounds has n_shards + 1
 * entries.  resident = 0 balances for the end-to-end path (link bound: bytes), 1 for resident batches (hashing + RSA).
 * Host-only. */
int zkb_plan_shards(const zkb_email_view *emails, size_t n, size_t n_shards, int resident, size_t *bounds);

/* (a) ONE process, several devices: one engine and one host thread per device behind a single handle. */
typedef struct zkb_multi zkb_multi;
typedef struct zkb_multi_regex zkb_multi_regex;
typedef struct zkb_multi_batch zkb_multi_batch;
/* opt->device is ignored; opt->host_threads (0 = all) is split evenly between the devices.  devices == NULL: 0..n-1. */
int zkb_multi_create(const zkb_options *opt, const int32_t *devices, size_t n_devices, zkb_multi **out);
void zkb_multi_destroy(zkb_multi *m);
size_t zkb_multi_devices(const zkb_multi *m);
zkb_engine *zkb_multi_engine(zkb_multi *m, size_t i);   /* the engine of device i (flags, statistics) */
int zkb_multi_host_register(zkb_multi *m, const void *p, size_t len);   /* page-locks once, usable by every device */
int zkb_multi_host_unregister(zkb_multi *m, const void *p);
int zkb_multi_regex_set_create(zkb_multi *m, const zkb_dfa_view *parts, size_t n_header, size_t n_body,
                               int header_present, int body_present, zkb_multi_regex **out);
void zkb_multi_regex_destroy(zkb_multi_regex *r);
/* zkb_verify_batch over all devices: every device's pipeline writes its range of `out`; no collective is needed for
 * the host to hold every result. */
int zkb_multi_verify_batch(zkb_multi *m, const zkb_email_view *emails, size_t n, const zkb_multi_regex *regex,
                           const zkb_email_captures *captures, zkb_result *out);
/* Resident form (raw != 0: zkb_batch_prepare_raw on every device). */
int zkb_multi_batch_prepare(zkb_multi *m, const zkb_email_view *emails, size_t n, const zkb_multi_regex *regex,
                            const zkb_email_captures *captures, int raw, zkb_multi_batch **out);
/* Runs every shard's kernels; gather != 0 adds the ncclAllGather of the result records so that every device holds the
 * whole batch's records.  Returns when all devices are done; *ms_max = slowest device (CUDA events). */
int zkb_multi_batch_run(zkb_multi_batch *mb, int gather, float *ms_max);
int zkb_multi_batch_fetch(zkb_multi_batch *mb, zkb_result *out);
int zkb_multi_batch_bounds(const zkb_multi_batch *mb, size_t *bounds, size_t n);   /* n_devices + 1 entries */
/* Copies the gathered records as held by device `device_index` to the host, in email order (rec_bytes each). */
int zkb_multi_batch_gathered(zkb_multi_batch *mb, size_t device_index, uint8_t *host_records, size_t *rec_bytes);
void zkb_multi_batch_destroy(zkb_multi_batch *mb);

/* (b) one process per device (torchrun-style): rank 0 obtains the id, the caller carries its 128 bytes to the other
 * processes by any means, every rank creates its communicator (collective call). */
#define ZKB_COMM_ID_BYTES 128
typedef struct zkb_comm zkb_comm;
int zkb_comm_unique_id(uint8_t id[ZKB_COMM_ID_BYTES]);
int zkb_comm_create(zkb_engine *e, const uint8_t id[ZKB_COMM_ID_BYTES], int rank, int world, zkb_comm **out);
void zkb_comm_destroy(zkb_comm *c);
/* Enqueues, on the engine stream behind the kernels of the batch's last zkb_batch_run*, the all-gather of its result
 * records.  *dev_records: world x *slot_records records of *rec_bytes bytes in this rank's device memory (rank r's
 * records start at r * slot_records; slot_records = the largest shard, shorter shards are zero padded). */
int zkb_comm_allgather_records(zkb_comm *c, zkb_batch *b, void **dev_records, size_t *slot_records, size_t *rec_bytes);
/* zkb_batch_run_async + the exchange of the records in one call: the records of resident chunk k travel (ncclSend / ncclRecv
 * on a stream of the communicator) while the kernels of chunk k+1 run; same output layout as zkb_comm_allgather_records,
 * complete in engine-stream order.  A collective: every rank of the communicator makes the call, and all ranks switch to a
 * newly prepared batch in the same call (the first call with a batch exchanges the ranks' chunk sizes). */
int zkb_comm_run_allgather(zkb_comm *c, zkb_batch *b, void **dev_records, size_t *slot_records, size_t *rec_bytes);
int zkb_comm_rank_records(const zkb_comm *c, uint64_t *counts, size_t n);   /* records held per rank */

/* ---- host-only utility: cfdkim::canonicalize_signed_email (core/src/circuits.rs:34-35,
 * helpers/src/generator.rs:63) ----
 * Returns the header-hash preimage and the canonical body of the first valid DKIM-Signature header
 * (the regex haystacks before the quoted-printable clean-up).  Needs no device (pure byte work, the
 * same code the batch path runs on its host threads).  *hdr / *body are malloc'ed; free with
 * zkb_free.  Returns ZKB_OK, or ZKB_E_INVALID when the reference's unwrap() would panic
 * (*detail = 1 parse error, 2 no valid signature, 3 bad c=, 4 bad l=, 5 bad b= base64). */
int zkb_host_canonicalize(const uint8_t *raw_email, size_t raw_email_len, int64_t now_unix,
                          uint8_t **hdr, size_t *hdr_len, uint8_t **body, size_t *body_len,
                          int *detail);

/* ---- host-only utility for the input generator (helpers/src/generator.rs:16-31): the DKIM-Signature headers of a
 * message, top to bottom, as generate_email_inputs walks them.  *out (malloc'ed, free with zkb_free) holds *n_sigs
 * records { u32 valid; u32 d_len; u32 s_len; d bytes; s bytes } back to back (little-endian, unpadded); valid = 1
 * when cfdkim::validate_header accepts the header (d / s are its d= and s= tag values, FWS removed), else 0 with
 * empty strings.  Returns ZKB_E_INVALID when mailparse::parse_mail would fail. */
int zkb_host_dkim_signatures(const uint8_t *raw_email, size_t raw_email_len, int64_t now_unix,
                             uint8_t **out, size_t *out_len, size_t *n_sigs);

/* ---- host-only: Solidity-ABI packing of the verifier outputs (core/src/io.rs:5-53 VerificationOutput::abi_encode,
 * helpers/src/io.rs:12-31 AbiDecodable; the step right after the hot path for on-chain consumers) ---- */
typedef struct zkb_str { const char *s; size_t len; } zkb_str;
typedef struct zkb_span { uint64_t off, len; } zkb_span;
/* One output to encode.  with_regex = 0: VerificationOutput::EmailOnly (SolEmailOutput, core/src/io.rs:6-10);
 * 1: VerificationOutput::WithRegex (SolEmailWithRegexOutput, core/src/io.rs:12-15).  external_inputs is the flattened
 * [name1, value1, ...] list of EmailVerifierOutput (core/src/circuits.rs:18-27). */
typedef struct zkb_output_view {
  const uint8_t *from_domain_hash;  /* 32 bytes */
  const uint8_t *public_key_hash;   /* 32 bytes */
  const zkb_str *external_inputs; size_t n_external_inputs;
  const zkb_str *matches; size_t n_matches;
  int with_regex;
} zkb_output_view;
typedef struct zkb_abi_decoded {
  int32_t with_regex;
  uint8_t from_domain_hash[32], public_key_hash[32];
  uint32_t n_external_inputs, n_matches;
} zkb_abi_decoded;
/* Encodes n outputs into one malloc'ed blob (free with zkb_free); output i occupies
 * [offsets[i], offsets[i+1]) — offsets has n+1 entries, caller-allocated.  threads <= 0: all host cores. */
int zkb_abi_encode_batch(const zkb_output_view *outs, size_t n, int threads, uint8_t **blob, uint64_t *offsets);
/* Decodes one blob the way AbiDecodable::abi_decode does (EmailOnly first, then WithRegex; the bytes must be the
 * canonical encoding).  *spans (malloc'ed, free with zkb_free) locates the n_external_inputs + n_matches strings
 * inside `data`.  Returns ZKB_E_INVALID where the reference returns Err. */
int zkb_abi_decode(const uint8_t *data, size_t len, zkb_abi_decoded *out, zkb_span **spans);

/* ---- kernel-level entry points (host buffers in, host buffers out) for parity tests ---- */
/* SHA-256 of n messages data[off[i] .. off[i]+len[i]) ; out = n x 32 bytes. */
int zkb_sha256_batch(zkb_engine *e, const uint8_t *data, size_t data_len, const uint64_t *off,
                     const uint32_t *len, size_t n, uint8_t *out);
/* RSA PKCS#1 v1.5 SHA-256 verify of n items: key DER, 32-byte digest, signature bytes.
 * ok[i] = 1 pass, 0 fail, 2 key rejected.  All keys may be distinct. */
int zkb_rsa_verify_batch(zkb_engine *e, const uint8_t *const *key_der, const size_t *key_len,
                         const uint8_t *digests, const uint8_t *const *sig, const size_t *sig_len,
                         size_t n, uint8_t *ok);
/* find_iter of one regex part over n haystacks; qp!=0 applies the soft-break cleaner on the fly.
 * out = n x (count, start, end, flags). */
int zkb_dfa_scan_batch(zkb_engine *e, const zkb_dfa_view *part, const uint8_t *data,
                       size_t data_len, const uint64_t *off, const uint32_t *len, size_t n, int qp,
                       uint32_t *out);
/* Integer-pipe peak microbenchmarks (register-only): giga thread-instructions per second for
 * [0]=IMAD.WIDE.U32(.X) (carry chains, 32x32+64 -> 64, the RSA inner instruction) [1]=IADD3 [2]=LOP3
 * [3]=SHF, then [4]=SM clock MHz (device attribute) [5]=SM count.  Used as the roofline denominators
 * of the integer-bound kernels (MEASURED_PEAKS.json has no integer peak). */
int zkb_int_pipe_peaks(zkb_engine *e, double out[8]);

/*
 * ZDF1 dense-DFA table layout (little-endian), the on-the-wire form of `DFA.fwd` / `DFA.bwd`:
 *   0   u32 magic 0x3146445A ("ZDF1")
 *   4   u32 flags: bit0 reverse, bit1 utf8 mode, bit2 pattern can match the empty string
 *   8   u32 n_states           (state 0 is the dead state)
 *   12  u32 n_classes          (byte classes + 1; the last class is end-of-input)
 *   16  u32 min_match, 20 u32 max_match   (match states form the id range [min,max]; they are
 *                                          entered one byte late, as in regex-automata)
 *   24  u32 start[12]          (unanchored then anchored; kinds NonWordByte, WordByte, Text,
 *                               LineLF, LineCR, CustomLineTerminator)
 *   72  u8  class_map[256]
 *   328 u8  start_map[256]     (look-behind byte -> start kind)
 *   584 u32 trans[n_states * n_classes]
 */
#define ZKB_ZDF_MAGIC 0x3146445Au
#define ZKB_ZDF_HEADER 584

#ifdef __cplusplus
}
#endif
#endif
