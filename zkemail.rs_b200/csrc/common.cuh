// common.cuh — shared device/host definitions for the zkemail_b200 kernels (sm_100a only).
#pragma once
#ifndef ZKB_HOST_EMU
#include <cuda_runtime.h>
#endif
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#define ZKB_CUDA_OK(expr)                                                                    \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      fprintf(stderr, "[zkemail_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e),   \
              __FILE__, __LINE__, cudaGetErrorString(_e));                                   \
      return ZKB_E_CUDA;                                                                     \
    }                                                                                        \
  } while (0)

// Key table entry layout (u32 words, little-endian limbs), one per unique public key.
//   [0..127]   n      (zero padded to 128 limbs)
//   [128..255] R^2 mod n, R = 2^(32*limbs)
//   [256] n0inv = -n^-1 mod 2^32   [257] e low 32   [258] e high   [259] k = byte length of n
//   [260] limbs class (32/64/128)   [261..263] reserved
#define ZKB_KEY_STRIDE 264
#define ZKB_KEY_RR 128
#define ZKB_KEY_N0INV 256
#define ZKB_KEY_ELO 257
#define ZKB_KEY_EHI 258
#define ZKB_KEY_K 259
#define ZKB_KEY_LIMBS 260

// One RSA work item (16 B, loaded as uint4 by every lane of the cooperating group).
//   x: word offset of the signature (LE limbs, zero padded to the key's limb class) in sig arena
//   y: key id   z: digest slot of the header hash   w: candidate slot (output)
struct __align__(16) RsaItem {
  uint32_t sig_off, key_id, digest_slot, cand;
};

// One DFA scan item: haystack = arena[hay_off .. hay_off + msg_len[msg]) (the length lives in the
// device-side message table because device-canonicalised bodies only get theirs on the device).
struct __align__(16) DfaItem {
  uint64_t hay_off;
  uint32_t msg;
  uint32_t out_slot;
};

// One body to canonicalise on the device (canon.cuh).
struct __align__(16) CanonItem {
  uint64_t raw_off;   // body bytes in the raw span buffer
  uint32_t raw_len;
  uint32_t msg;       // message index: arena slot msg_off[msg], length written to msg_len[msg]
  uint32_t flags;     // bit0 relaxed, bit1 has l=
  uint32_t l;         // l= value (clamped to 2^32-1)
  uint32_t pad[2];
};

// ---- device front end (frontend.cuh) records ----
#define FE_MAXH 64      // headers per message handled on the device
#define FE_MAXN 16      // names in h=
#define FE_PRE_CAP 4096 // bytes reserved per header preimage slot

enum { FE_FALLBACK = 1u, FE_MAIL_PARSE = 2u, FE_BH_VALID = 4u, FE_SIG_SYNTAX = 8u, FE_SIG_BADLEN = 16u,
       FE_HDR_RELAXED = 32u, FE_BODY_RELAXED = 64u,
       FE_MULTI = 128u /* several DKIM-Signature headers: a candidate that fails goes to the host front end */,
       FE_HAS_L = 256u /* l= present: the canonical body is truncated (CanonItem.l) */ };

struct __align__(16) FeIn {   // host-built, one per message
  uint64_t raw_off;           // message bytes in the raw span buffer
  uint32_t raw_len;
  uint32_t dom_off;           // from_domain bytes in the arena (staged once per distinct domain)
  uint32_t dom_len;
  uint32_t k;                 // byte length of the key's modulus
  uint32_t limbs;             // limb class of the key (32 / 64 / 128)
  uint32_t sig_word_off;      // where this message's signature limbs go (words)
  uint32_t body_msg, pre_msg; // message-table indices of the canonical body / header preimage slots
  uint32_t cand;              // signature-candidate index (bh= words go to cand_bh[8*cand..])
  uint32_t email;             // chunk-local email index (result record slot, assemble.cuh)
  uint32_t dom_msg, key_msg;  // message-table indices of the from_domain / key hashes
  uint32_t pad[2];
};
struct __align__(16) FeOut {
  uint32_t flags;
  uint32_t body_off, body_len;   // raw body inside the message (for host-side capture checks)
  uint32_t pre_len;
  uint32_t bh[8];
};

// Device-built result records (assemble.cuh): the head of zkb_result followed by one uint4 per regex part
#define ZKB_REC_HEAD 144            // offsetof(zkb_result, parts)
#define ZKB_REC_REDO 0x7fffffff     // status of a message the device declined: the host front end decides

// Per-candidate flags written by the device
#define ZKB_F_RSA_OK 1u
#define ZKB_F_BH_OK 2u
