// assemble.cuh — per-email result records built on the device (sm_100a).
//
// Replaces, for messages the device front end handled, the host-side control flow of zkemail_core::verify_email /
// verify_email_with_regex (core/src/circuits.rs:9-29, 31-68): DKIM verdict -> assert!(verified) -> the two output
// hashes -> per regex part "exactly one match" (core/src/regex.rs:36-39) in order, stopping at the first failing part.
// One thread per signature candidate reads what the earlier kernels left in HBM (front-end flags, bh= / RSA flags,
// SHA-256 state words, DFA scan results) and writes one fixed-size record: the first 144 bytes of zkb_result
// (status .. n_parts) followed by P part entries.  The records are what travels device -> host (one D2H per chunk,
// copied straight into the caller's array) and what the multi-GPU all-gather exchanges (verdict + hashes + spans).
// Expected-capture substring checks need the caller's strings and stay on the host (they can only turn a pass of a
// part into a fail).  Messages the device declined carry the status ZKB_REC_REDO and are re-run by the host front end.
#pragma once
#include "../../include/zkemail_b200.h"
#include "common.cuh"

namespace zkb {

__device__ __forceinline__ uint32_t asm_be(uint32_t w) { return __byte_perm(w, 0, 0x0123); }

// rec_words: record stride in 32-bit words (36 + 4 * P rounded up to a multiple of 4)
__global__ void __launch_bounds__(128)
assemble_kernel(const FeIn* __restrict__ in, const FeOut* __restrict__ fo_arr, uint32_t n, const uint32_t* __restrict__ cand_flags,
                const uint32_t* __restrict__ digests, const uint4* __restrict__ dfa_out, uint32_t P, uint32_t body_mask,
                uint32_t have_regex, uint32_t* __restrict__ recs, uint32_t rec_words) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const FeIn fi = in[idx];
  const uint32_t fl = fo_arr[idx].flags;
  uint32_t* r = recs + (size_t)fi.email * rec_words;
  uint32_t w[36];
#pragma unroll
  for (int i = 0; i < 36; i++) w[i] = 0;
  int32_t status = ZKB_ST_OK, detail = ZKB_DKIM_NEUTRAL;
  uint32_t n_parts = 0;
  bool hashes = false, pass = false;
  if (fl & FE_MAIL_PARSE) status = ZKB_ST_MAIL_PARSE;
  else if (fl & FE_FALLBACK) status = ZKB_REC_REDO;
  else {
    const uint32_t f = cand_flags[fi.cand];
    hashes = true;
    const bool bh_ok = (fl & FE_BH_VALID) && (f & ZKB_F_BH_OK);
    detail = ZKB_DKIM_PASS;
    if (!bh_ok) detail = ZKB_DKIM_BODY_HASH;
    else if (fl & FE_SIG_SYNTAX) detail = ZKB_DKIM_SIG_SYNTAX;
    else if ((fl & FE_SIG_BADLEN) || !(f & ZKB_F_RSA_OK)) detail = ZKB_DKIM_SIG_MISMATCH;
    // several signature headers: the reference goes on to the later ones; only a pass is final on the device
    if (detail != ZKB_DKIM_PASS && (fl & FE_MULTI)) status = ZKB_REC_REDO;
    else if (detail != ZKB_DKIM_PASS) status = ZKB_ST_DKIM_FAIL;
    else pass = true;
    w[34] = (bh_ok ? 1u : 0u) | (pass ? 0x100u : 0u);   // bh_ok, rsa_ok, pad[2]
  }
  if (hashes) {
    const uint32_t* db = digests + (size_t)fi.body_msg * 8;
    const uint32_t* dh = digests + (size_t)fi.pre_msg * 8;
#pragma unroll
    for (int i = 0; i < 8; i++) { w[2 + i] = asm_be(db[i]); w[10 + i] = asm_be(dh[i]); }
  }
  if (pass) {
    const uint32_t* dd = digests + (size_t)fi.dom_msg * 8;
    const uint32_t* dk = digests + (size_t)fi.key_msg * 8;
#pragma unroll
    for (int i = 0; i < 8; i++) { w[18 + i] = asm_be(dd[i]); w[26 + i] = asm_be(dk[i]); }
  }
  // regex parts in order; the first part without exactly one match ends the evaluation (its entry is kept)
  uint4* rp = reinterpret_cast<uint4*>(r + 36);
  if (pass && have_regex) {
    const uint4* d = dfa_out + (size_t)fi.email * P;
    bool alive = true;
    for (uint32_t p = 0; p < P; p++) {
      uint4 e = make_uint4(0u, 0u, 0u, 0u);
      if (alive) {
        const uint4 v = d[p];
        const bool ok = v.x == 1u;
        e = make_uint4(v.x, v.y, v.z, ok ? 1u : 0u);
        n_parts = p + 1;
        if (!ok) { status = ((body_mask >> p) & 1u) ? ZKB_ST_REGEX_BODY : ZKB_ST_REGEX_HEADER; alive = false; }
      }
      rp[p] = e;
    }
  } else {
    for (uint32_t p = 0; p < P; p++) rp[p] = make_uint4(0u, 0u, 0u, 0u);
  }
  w[0] = (uint32_t)status; w[1] = (uint32_t)detail; w[35] = n_parts;
  uint4* r4 = reinterpret_cast<uint4*>(r);
#pragma unroll
  for (int i = 0; i < 9; i++) r4[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

}  // namespace zkb
