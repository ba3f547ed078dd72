// sha256.cuh — multi-message SHA-256 for sm_100a (kernel K1 of DESIGN.md).
//
// Replaces, on the device, every sha2::Sha256 call on the reference hot path:
//   hash_bytes                (core/src/crypto.rs:3-7; circuits.rs:16-17: from_domain, key DER)
//   cfdkim body hash / header hash (called from core/src/email.rs:31-33)
//
// Mapping: lane = message; a warp advances 32 messages in lock step (the host orders the message
// list by block count so lanes of a warp stay converged).  Each lane streams its own message
// with 128-bit loads: a 64-byte block is 4 x LDG.128 out of two 32-byte sectors, every fetched
// byte is used, and the 126 MB L2 / 228 KB L1 absorb the line sharing between consecutive loads.
// Padding (0x80, zeros, 64-bit length) is generated in registers, so the arena holds raw bytes.
// The compression function is fully unrolled with a 16-word ring for the message schedule;
// the bound is the INT32 ALU pipe (SHF/LOP3/IADD3), not HBM.
#pragma once
#include "common.cuh"

namespace zkb {

__device__ __constant__ uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

#ifndef ZKB_SHA_ROT_DEFAULT
#define ZKB_SHA_ROT_DEFAULT 0
#endif
__device__ __forceinline__ uint32_t rotr32(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// 32-bit add issued on the FMA pipe (IMAD: a * one + b, `one` is a run-time 1 the compiler cannot fold).
// SHA-256 is bound by the ALU pipe (SHF / LOP3 / IADD3, 64 lanes/clk/SM); the FMA pipe is idle, so the
// additions are moved there and the ALU pipe keeps only the rotates and the boolean functions.
__device__ __forceinline__ uint32_t fadd(uint32_t a, uint32_t b, uint32_t one) {
#ifdef ZKB_HOST_EMU
  return a * one + b;
#else
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
  return d;
#endif
}

// A rotate issued on the FMA pipe: x * 2^(32-k) is the 64-bit value (x >> k) : (x << (32-k)) - one IMAD.WIDE - and the two
// disjoint halves are merged with an IMAD add.  pw = one << (32 - k), built from the run-time 1 so that the compiler keeps
// the multiplication.  3 FMA-pipe slots buy one ALU-pipe slot: worth it for as many rotates as it takes to level the pipes.
__device__ __forceinline__ uint32_t rotr_fma(uint32_t x, uint32_t pw, uint32_t one) {
#ifdef ZKB_HOST_EMU
  const uint64_t p = (uint64_t)x * pw;
  return (uint32_t)(p >> 32) * one + (uint32_t)p;
#else
  uint32_t d;
  asm("{\n\t.reg .u64 p;\n\t.reg .u32 lo, hi;\n\tmul.wide.u32 p, %1, %2;\n\tmov.b64 {lo, hi}, p;\n\tmad.lo.u32 %0, hi, %3, lo;\n\t}"
      : "=r"(d) : "r"(x), "r"(pw), "r"(one));
  return d;
#endif
}

// One compression; w[16] holds the big-endian message words and is clobbered.  ROT: how many of the rotate families go to
// the FMA pipe (0 none; 1: rotr 7 of the schedule, 48 per block; 2: + rotr 17, 96 per block; 3: + rotr 6 of the rounds, 160).
template <int ROT = ZKB_SHA_ROT_DEFAULT>
__device__ __forceinline__ void sha256_compress(uint32_t st[8], uint32_t w[16], uint32_t one) {
  uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
  const uint32_t pw7 = one << 25, pw17 = one << 15, pw6 = one << 26;
#pragma unroll
  for (int i = 0; i < 64; i++) {
    if (i >= 16) {
      uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
      uint32_t s0 = (ROT >= 1 ? rotr_fma(w15, pw7, one) : rotr32(w15, 7)) ^ rotr32(w15, 18) ^ (w15 >> 3);
      uint32_t s1 = (ROT >= 2 ? rotr_fma(w2, pw17, one) : rotr32(w2, 17)) ^ rotr32(w2, 19) ^ (w2 >> 10);
      w[i & 15] = fadd(fadd(w[i & 15], s0, one), fadd(w[(i - 7) & 15], s1, one), one);
    }
    uint32_t S1 = (ROT >= 3 ? rotr_fma(e, pw6, one) : rotr32(e, 6)) ^ rotr32(e, 11) ^ rotr32(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);  // one LOP3
    uint32_t t1 = fadd(fadd(fadd(h, S1, one), ch, one), fadd(w[i & 15], SHA_K[i], one), one);
    uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);  // one LOP3
    h = g; g = f; f = e; e = fadd(d, t1, one); d = c; c = b; b = a; a = fadd(fadd(t1, S0, one), mj, one);
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// msg_off: byte offsets into arena, 16-byte aligned; every message slot must be readable up to
// the next 64-byte boundary past its end (the packer guarantees it).  order: optional
// permutation (message ids sorted by block count).  digests: n x 8 native state words.
// PREFETCH: the 64 bytes of block b+1 are requested before block b is compressed.  Used when the launch has few
// lanes (large bodies): with ~5 warps per scheduler the load latency at the top of every block is not hidden by
// other warps.  Costs 16 registers, so launches with many lanes (occupancy-bound) use the plain variant.
template <bool PREFETCH, int ROT = ZKB_SHA_ROT_DEFAULT>
__global__ void __launch_bounds__(128)
sha256_batch_kernel(const uint8_t* __restrict__ arena, const uint64_t* __restrict__ msg_off,
                    const uint32_t* __restrict__ msg_len, const uint32_t* __restrict__ order,
                    uint32_t n, uint32_t* __restrict__ digests, uint32_t one) {
  uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  uint32_t m = order ? order[idx] : idx;
  const uint4* p = reinterpret_cast<const uint4*>(arena + msg_off[m]);
  uint32_t len = msg_len[m];
  uint32_t nfull = len >> 6, rem = len & 63;
  uint32_t total = nfull + 1 + (rem >= 56 ? 1 : 0);
  uint32_t st[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                    0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  uint4 n0, n1, n2, n3;   // PREFETCH: the next block's bytes (block 0 first; the slot is readable past nfull)
  if (PREFETCH) { n0 = __ldg(p); n1 = __ldg(p + 1); n2 = __ldg(p + 2); n3 = __ldg(p + 3); }
  for (uint32_t blk = 0; blk < total; blk++) {
    uint32_t w[16];
    if (blk <= nfull) {
      uint4 v0, v1, v2, v3;
      if (PREFETCH) {
        v0 = n0; v1 = n1; v2 = n2; v3 = n3;
      } else {
        v0 = __ldg(p + 4 * blk); v1 = __ldg(p + 4 * blk + 1); v2 = __ldg(p + 4 * blk + 2); v3 = __ldg(p + 4 * blk + 3);
      }
      w[0] = bswap32(v0.x); w[1] = bswap32(v0.y); w[2] = bswap32(v0.z); w[3] = bswap32(v0.w);
      w[4] = bswap32(v1.x); w[5] = bswap32(v1.y); w[6] = bswap32(v1.z); w[7] = bswap32(v1.w);
      w[8] = bswap32(v2.x); w[9] = bswap32(v2.y); w[10] = bswap32(v2.z); w[11] = bswap32(v2.w);
      w[12] = bswap32(v3.x); w[13] = bswap32(v3.y); w[14] = bswap32(v3.z); w[15] = bswap32(v3.w);
    }
    if (PREFETCH) {
      // unconditional (a predicated load made ptxas copy the sixteen registers around it and wait for the load at once):
      // past the last readable block the index is clamped and the bytes are ignored
      const uint32_t nx = blk < nfull ? blk + 1 : nfull;
      n0 = __ldg(p + 4 * nx); n1 = __ldg(p + 4 * nx + 1); n2 = __ldg(p + 4 * nx + 2); n3 = __ldg(p + 4 * nx + 3);
    }
    if (blk >= nfull) {  // tail: mask bytes past the end, append 0x80 / zeros / bit length
#pragma unroll
      for (int i = 0; i < 16; i++) {
        int r = (blk == nfull) ? (int)rem - 4 * i : -1;  // message bytes remaining in this word
        uint32_t v = (blk == nfull) ? w[i] : 0u;
        if (r >= 4) {
        } else if (r > 0) v = (v & (0xFFFFFFFFu << (32 - 8 * r))) | (0x80u << (24 - 8 * r));
        else if (r == 0) v = 0x80000000u;
        else v = 0u;
        w[i] = v;
      }
      if (blk == total - 1) { w[14] = len >> 29; w[15] = len << 3; }
    }
    sha256_compress<ROT>(st, w, one);
  }
  uint4* o = reinterpret_cast<uint4*>(digests + (size_t)m * 8);
  o[0] = make_uint4(st[0], st[1], st[2], st[3]);
  o[1] = make_uint4(st[4], st[5], st[6], st[7]);
}

// bh= check: body digest of the candidate vs the decoded bh= value (stored as 8 native words).
__global__ void bh_check_kernel(const uint32_t* __restrict__ digests,
                                const uint32_t* __restrict__ body_slot,
                                const uint32_t* __restrict__ bh_words, uint32_t n_cand,
                                uint32_t* __restrict__ cand_flags) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cand) return;
  const uint32_t* d = digests + (size_t)body_slot[c] * 8;
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 8; i++) ok = ok && (d[i] == bh_words[(size_t)c * 8 + i]);
  if (ok) atomicOr(cand_flags + c, ZKB_F_BH_OK);
}

}  // namespace zkb
