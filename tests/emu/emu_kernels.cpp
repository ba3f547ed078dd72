// emu_kernels.cpp — TEST INFRASTRUCTURE: builds the product's kernel sources for the CPU through
// tests/emu/emu_cuda.h and exposes them with a C ABI so the `not gpu` tests can check the very
// source the GPU runs (SHA-256 lanes, cooperative Montgomery groups, DFA scans) against the
// oracle, hashlib and Python big ints.
#include "emu_cuda.h"

#include "../../include/zkemail_b200.h"
#include "../../zkemail.rs_b200/csrc/common.cuh"
#include "../../zkemail.rs_b200/csrc/keytab.hpp"
#include "../../zkemail.rs_b200/csrc/sha256.cuh"
#include "../../zkemail.rs_b200/csrc/rsa.cuh"
#include "../../zkemail.rs_b200/csrc/canon.cuh"
#include "../../zkemail.rs_b200/csrc/dkim_host.hpp"
#include "../../zkemail.rs_b200/csrc/frontend.cuh"
#include "../../zkemail.rs_b200/csrc/frontend_warp.cuh"
// dfa.cuh added below once rewritten

using namespace zkb;

template <int LIMBS, int T, bool G>
static void run_rsa(const uint32_t* sig_arena, const RsaItem* items, uint32_t n_items,
                    const uint32_t* keytab, const uint32_t* digests, uint32_t* flags) {
  const unsigned block = 128;
  unsigned threads = n_items * T;
  emu::launch((threads + block - 1) / block, block, [&]() {
    rsa_verify_kernel<LIMBS, T, G>(sig_arena, items, n_items, keytab, digests, flags);
  });
}

template <int V>
static void run_rsa_sqr(const uint32_t* sig_arena, const RsaItem* items, uint32_t n_items,
                        const uint32_t* keytab, const uint32_t* digests, uint32_t* flags) {
  const unsigned block = 64;
  unsigned threads = n_items * 4;
  emu::launch((threads + block - 1) / block, block, [&]() {
    rsa_verify_kernel<64, 4, false, V>(sig_arena, items, n_items, keytab, digests, flags);
  });
}

extern "C" {

void emu_sha256_batch(const uint8_t* arena, const uint64_t* off, const uint32_t* len,
                      const uint32_t* order, uint32_t n, uint32_t* digests) {
  const unsigned block = 128;
  emu::launch((n + block - 1) / block, block,
              [&]() { sha256_batch_kernel<false>(arena, off, len, order, n, digests, 1u); });
}
// the variant that requests block b+1 before compressing block b (launches with few lanes)
void emu_sha256_batch_prefetch(const uint8_t* arena, const uint64_t* off, const uint32_t* len,
                               const uint32_t* order, uint32_t n, uint32_t* digests) {
  const unsigned block = 128;
  emu::launch((n + block - 1) / block, block,
              [&]() { sha256_batch_kernel<true>(arena, off, len, order, n, digests, 1u); });
}

// the instantiations that issue rotates on the FMA pipe (rot = 1 .. 3 rotate families as multiply + add)
void emu_sha256_batch_rot(const uint8_t* arena, const uint64_t* off, const uint32_t* len,
                          const uint32_t* order, uint32_t n, uint32_t* digests, int rot) {
  const unsigned block = 128;
  emu::launch((n + block - 1) / block, block, [&]() {
    if (rot == 1) sha256_batch_kernel<false, 1>(arena, off, len, order, n, digests, 1u);
    else if (rot == 2) sha256_batch_kernel<true, 2>(arena, off, len, order, n, digests, 1u);
    else sha256_batch_kernel<false, 3>(arena, off, len, order, n, digests, 1u);
  });
}

// key DER -> key-table entry (ZKB_KEY_STRIDE words). returns 0 ok, 1 rejected
int emu_build_key_entry(const uint8_t* der, size_t len, uint32_t* ent) {
  RsaKeyInfo k;
  if (!parse_rsa_public_key(der, len, k)) return 1;
  build_key_entry(k, ent);
  return 0;
}

int emu_rsa_verify(int limbs, int T, int generic, const uint32_t* sig_arena, const void* items,
                   uint32_t n_items, const uint32_t* keytab, const uint32_t* digests,
                   uint32_t* flags) {
  const RsaItem* it = (const RsaItem*)items;
  if (limbs == 64 && T == 104 && !generic) { run_rsa_sqr<8>(sig_arena, it, n_items, keytab, digests, flags); return 0; }   // T = 104: the squaring variant
  if (limbs == 64 && T == 204 && !generic) { run_rsa_sqr<104>(sig_arena, it, n_items, keytab, digests, flags); return 0; }  // T = 204: with the rolled combine
#define CASE(LB, TT)                                                                        \
  if (limbs == LB && T == TT) {                                                             \
    if (generic) run_rsa<LB, TT, true>(sig_arena, it, n_items, keytab, digests, flags);     \
    else run_rsa<LB, TT, false>(sig_arena, it, n_items, keytab, digests, flags);            \
    return 0;                                                                               \
  }
  CASE(32, 2) CASE(32, 4) CASE(32, 8) CASE(32, 16)
  CASE(64, 2) CASE(64, 4) CASE(64, 8) CASE(64, 16) CASE(64, 32)
  CASE(128, 8) CASE(128, 16) CASE(128, 32)
#undef CASE
  return 1;
}

}  // extern "C"

// ---- regex compiler + DFA kernel ----
#include "../../zkemail.rs_b200/csrc/regexc.hpp"
#include "../../zkemail.rs_b200/csrc/dfa_host.hpp"

extern "C" {

// returns 0 ok; *fwd/*bwd malloc'ed (free with emu_free)
int emu_regex_compile(const char* pat, size_t len, uint8_t** fwd, size_t* fl, uint8_t** bwd, size_t* bl,
                      char* err, size_t err_cap) {
  std::vector<uint8_t> f, b;
  std::string e;
  if (!rx::compile(pat, len, f, b, e)) {
    if (err && err_cap) { strncpy(err, e.c_str(), err_cap - 1); err[err_cap - 1] = 0; }
    return 1;
  }
  *fwd = (uint8_t*)malloc(f.size()); memcpy(*fwd, f.data(), f.size()); *fl = f.size();
  *bwd = (uint8_t*)malloc(b.size()); memcpy(*bwd, b.data(), b.size()); *bl = b.size();
  return 0;
}
void emu_free(void* p) { free(p); }

// out = n x uint4 (count, start, end, panic).  table_form: 0 = as the engine chooses, 1 = force
// class-compressed u16/u32, 2 = force u32 elements
int emu_dfa_scan(const uint8_t* fwd, size_t fl, const uint8_t* bwd, size_t bl, const uint8_t* arena,
                 const uint64_t* off, const uint32_t* len, uint32_t n, int qp, int use_smem, uint32_t* out, int table_form) {
  std::vector<uint8_t> fb, rb;
  uint32_t elem = 2;
  bool direct = false;
  if (table_form == 0) {
    if (!build_dfa_pair(fwd, fl, bwd, bl, fb, rb, elem, direct)) return 1;
  } else {
    uint32_t e1, e2;
    elem = table_form == 2 ? 4 : 0;
    if (!build_dfa_blob(fwd, fl, false, false, elem, fb, e1) || !build_dfa_blob(bwd, bl, true, false, elem ? elem : e1, rb, e2)) return 1;
    if (e1 != e2) {  // re-encode both as u32
      if (!build_dfa_blob(fwd, fl, false, false, 4, fb, e1) || !build_dfa_blob(bwd, bl, true, false, 4, rb, e2)) return 1;
    }
    elem = e1;
  }
  std::vector<DfaItem> items(n);
  for (uint32_t i = 0; i < n; i++) { items[i].hay_off = off[i]; items[i].msg = i; items[i].out_slot = i; }
  const unsigned block = 128;
  if (fb.size() + rb.size() > 200000) use_smem = 0;   // as the launchers do: tables that do not fit stay in global memory
  emu::launch((n + block - 1) / block, block, [&]() {
#define RUN1(TT, D, S) dfa_scan_kernel<TT, D, S>(arena, items.data(), n, len, fb.data(), (uint32_t)fb.size(), rb.data(), (uint32_t)rb.size(), qp, (uint4*)out)
#define RUN(TT, D) do { if (use_smem) RUN1(TT, D, true); else RUN1(TT, D, false); } while (0)
    if (elem == 2) { if (direct) RUN(uint16_t, true); else RUN(uint16_t, false); }
    else { if (direct) RUN(uint32_t, true); else RUN(uint32_t, false); }
#undef RUN1
#undef RUN
  });
  return 0;
}

// device-side body canonicalisation: n bodies at span[off[i] .. off[i]+len[i]); out slots of cap[i] bytes
// (64-byte aligned) are packed into `arena`; out_len[i] receives the canonical lengths
void emu_canon_body(const uint8_t* span, const uint64_t* off, const uint32_t* len, const uint32_t* flags, const uint32_t* lval,
                    uint32_t n, uint8_t* arena, const uint64_t* slot_off, uint32_t* out_len) {
  std::vector<CanonItem> items(n);
  for (uint32_t i = 0; i < n; i++) {
    items[i].raw_off = off[i]; items[i].raw_len = len[i]; items[i].msg = i; items[i].flags = flags[i]; items[i].l = lval[i];
  }
  const unsigned block = 128;
  if (getenv("ZKB_EMU_CANON_UNSTAGED")) emu::launch((n + block - 1) / block, block, [&]() { canon_body_kernel(span, items.data(), n, arena, slot_off, out_len); });
  else emu::launch((n + block - 1) / block, block, [&]() { canon_body_staged_kernel(span, items.data(), n, arena, slot_off, out_len, nullptr, nullptr, n); });
}

// What the device front end produced for one message, checked against the host front end (dkim_host.hpp).
// returns 0 = the device path declines (fallback), 1 = live and identical to the host, 2 = both report a
// mail parse error, negative = MISMATCH (code tells which field).
static int fe_check_against_host(const uint8_t* raw, uint32_t n, const uint8_t* dom, uint32_t dom_len, uint32_t k, uint32_t limbs, int allow_skip,
                                 const FeOut& fo, const std::vector<uint8_t>& pre, const std::vector<uint32_t>& sigw, uint32_t body_l) {
  std::vector<HeaderField> hs;
  size_t body_off = 0;
  const bool parsed = parse_headers(raw, n, hs, body_off);
  if (fo.flags & FE_MAIL_PARSE) return parsed ? -1 : 2;
  if (fo.flags & FE_FALLBACK) return 0;
  if (!parsed) return -2;
  // live on the device: the host must reach the cryptographic checks with the same bytes
  DkimSig sig;
  std::string scratch;
  int n_sig = 0, idx = -1;
  // the reference's walk over the DKIM-Signature headers (cfdkim::verify_email_with_key): headers of another domain
  // are skipped; the device may only have gone past headers that validate, and must have stopped at the first
  // header of `dom`
  for (size_t i = 0; i < hs.size(); i++) {
    if (!ieq_ascii(raw + hs[i].key_off, hs[i].key_len, "DKIM-Signature", 14)) continue;
    n_sig++;
    if (idx >= 0) continue;
    if (validate_dkim_header(raw + hs[i].val_off, hs[i].val_len, 1, sig) != ZKB_DKIM_PASS) return -4;
    const Tag* tdd = sig.get("d");
    if (ieq_ascii(sig.val(tdd), tdd->val_len, (const char*)dom, dom_len)) idx = (int)i;
    else if (!allow_skip) return -3;
  }
  if (idx < 0) return -5;
  if (((fo.flags & FE_MULTI) != 0) != (n_sig > 1)) return -18;
  bool hr, br;
  if (!parse_canon_tag(sig, hr, br)) return -6;
  if (!sig.val_is(sig.get("a"), "rsa-sha256")) return -7;
  {
    uint64_t l = 0;
    const Tag* tl = sig.get("l");
    if (tl && !parse_usize_tag(sig, tl, l)) return -8;   // the device must have declined an l= the reference rejects
    if ((tl != nullptr) != ((fo.flags & FE_HAS_L) != 0)) return -19;
    if (tl && body_l != (l > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)l)) return -20;
  }
  if (hr != ((fo.flags & FE_HDR_RELAXED) != 0) || br != ((fo.flags & FE_BODY_RELAXED) != 0)) return -9;
  size_t bl = 0;
  const uint8_t* b = find_body(raw, n, bl, body_off);
  if ((size_t)(b - raw) != fo.body_off || bl != fo.body_len) return -10;
  std::vector<uint8_t> hp(preimage_bound(body_off, sig.n));
  size_t pl = build_header_preimage(raw, hs, sig, hr, hp.data(), scratch);
  if (pl != fo.pre_len || memcmp(hp.data(), pre.data(), pl) != 0) {
    if (getenv("ZKB_EMU_DEBUG")) {
      fprintf(stderr, "host preimage (%zu): ", pl); fwrite(hp.data(), 1, pl, stderr);
      fprintf(stderr, "\ndevice preimage (%u): ", fo.pre_len); fwrite(pre.data(), 1, fo.pre_len, stderr); fprintf(stderr, "\n");
    }
    return -11;
  }
  const Tag* tbh = sig.get("bh");
  uint8_t bh[48];
  bool bh_valid = tbh->val_len == 44 && base64_decode(sig.val(tbh), 44, bh) == 32;
  if (bh_valid != ((fo.flags & FE_BH_VALID) != 0)) return -12;
  if (bh_valid)
    for (int i = 0; i < 8; i++)
      if (fo.bh[i] != (((uint32_t)bh[4 * i] << 24) | ((uint32_t)bh[4 * i + 1] << 16) | ((uint32_t)bh[4 * i + 2] << 8) | bh[4 * i + 3])) return -13;
  const Tag* tb = sig.get("b");
  std::vector<uint8_t> tmp(tb->val_len + 4);
  long sl = base64_decode(sig.val(tb), tb->val_len, tmp.data());
  if ((sl < 0) != ((fo.flags & FE_SIG_SYNTAX) != 0)) return -14;
  if (sl >= 0 && (((size_t)sl != k) != ((fo.flags & FE_SIG_BADLEN) != 0))) return -15;
  if (sl >= 0 && (size_t)sl == k) {
    std::vector<uint32_t> w(limbs, 0);
    for (long i = 0; i < sl; i++) { long bi = sl - 1 - i; w[bi >> 2] |= (uint32_t)tmp[i] << (8 * (bi & 3)); }
    if (w != sigw) return -16;
  } else {
    for (uint32_t x : sigw) if (x != 0) return -17;
  }
  return 1;
}

// the device reads whole aligned 16-byte blocks: give it the message at an odd offset inside a padded buffer
static uint8_t* fe_place(std::vector<uint8_t>& padded, const uint8_t* raw_in, uint32_t n) {
  padded.assign((size_t)n + 96, 0x3B);
  uint8_t* raw = padded.data() + 16 + ((16 - ((uintptr_t)padded.data() & 15)) & 15) + 5;
  memcpy(raw, raw_in, n);
  return raw;
}

// Scalar device front end (frontend.cuh: fe_process) against the host front end on one message.
int emu_fe_compare(const uint8_t* raw_in, uint32_t n, const uint8_t* dom, uint32_t dom_len, uint32_t k, uint32_t limbs, int allow_skip) {
  std::vector<uint8_t> padded;
  uint8_t* raw = fe_place(padded, raw_in, n);
  std::vector<uint8_t> pre(FE_PRE_CAP + 64, 0xEE);
  std::vector<uint32_t> sigw(limbs, 0xDEADBEEFu);
  FeOut fo;
  uint32_t body_l = 0;
  fe_process(raw, n, dom, dom_len, k, limbs, pre.data(), sigw.data(), fo, body_l, allow_skip != 0, 1);
  return fe_check_against_host(raw, n, dom, dom_len, k, limbs, allow_skip, fo, pre, sigw, body_l);
}

// Warp-cooperative device front end (frontend_warp.cuh: fe_process_warp, 32 emulated lanes) against the host front
// end.  *scalar_rc receives the scalar twin's verdict on the same message (the warp form may decline more, never less
// exactly: a message it accepts must compare equal to the host).
int emu_fe_compare_warp(const uint8_t* raw_in, uint32_t n, const uint8_t* dom, uint32_t dom_len, uint32_t k, uint32_t limbs, int allow_skip,
                        int* scalar_rc) {
  std::vector<uint8_t> padded;
  uint8_t* raw = fe_place(padded, raw_in, n);
  std::vector<uint8_t> pre(FE_PRE_CAP + 64, 0xEE);
  std::vector<uint32_t> sigw(limbs, 0xDEADBEEFu);
  FeOut fo;
  memset(&fo, 0, sizeof fo);
  uint32_t body_l = 0;
  static FeLut lut;
  fe_lut_fill(&lut, 0, 1);
  emu::launch(1, 32, [&]() {
    static FeWarpSmem sm;
    FeOut mine;
    uint32_t bl = 0;
    fe_process_warp(&sm, &lut, raw, n, dom, dom_len, k, limbs, pre.data(), sigw.data(), mine, bl, allow_skip != 0, 1);
    if ((threadIdx.x & 31) == 0) { fo = mine; body_l = bl; }
  });
  if (scalar_rc) *scalar_rc = emu_fe_compare(raw_in, n, dom, dom_len, k, limbs, allow_skip);
  return fe_check_against_host(raw, n, dom, dom_len, k, limbs, allow_skip, fo, pre, sigw, body_l);
}

}  // extern "C"
