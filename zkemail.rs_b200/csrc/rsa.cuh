// rsa.cuh — batched RSA PKCS#1 v1.5 (SHA-256) signature verification for sm_100a (kernel K2).
//
// Replaces rsa 0.9.6 `RsaPublicKey::verify(Pkcs1v15Sign::new::<Sha256>(), hash, sig)` as reached
// from cfdkim::verify_email_with_key (core/src/email.rs:31-33; SURVEY.md A.2 "RSA verify"):
//   reject unless len(sig)==k (host) and s<n;  em = s^e mod n;  accept iff
//   em == 00 01 FF.. 00 || 3031300d060960864801650304020105000420 || hash.
//
// Arithmetic: Montgomery multiplication (CIOS) on 32-bit limbs with R = 2^(32*LIMBS).
// One signature is handled by T cooperating lanes of a warp, each holding L = LIMBS/T limbs of
// every operand in registers.  Per multiplier limb b_i (warp shuffle broadcast) every lane does
// 2L multiply-accumulates: L for a*b_i and L for n*m.  Products are accumulated into two
// register arrays, one aligned to even and one to odd limb offsets, so that every 32x32->64
// product is ONE IMAD.WIDE.U32(.X) with predicate carry (ptxas fuses each
// mad.lo.cc/madc.hi.cc pair).  The division by 2^32 of each CIOS step is free: the 3-operand
// multiply-add reads its addend one register pair higher than it writes, and the two arrays swap
// roles.  Between lanes the low limb of lane p+1 moves to the top of lane p by one shuffle per
// step; carries between lanes are deferred in two spare limbs per array and resolved once per
// multiplication with a ballot-based generate/propagate scan.  tools/mont_model.py is the
// instruction-level model this code was transcribed from.
//
// e = 65537 fast path: to-Montgomery (x*R^2), 16 squarings, final multiply by plain s (leaves
// Montgomery form) = 18 multiplications = 18*(2*LIMBS^2) wide MACs.  Other exponents take the
// generic left-to-right ladder in a separate instantiation.
// Bound: the FMA pipe (IMAD.WIDE), not HBM and not tensor cores.
#pragma once
#include "common.cuh"

namespace zkb {

#ifndef ZKB_HOST_EMU
// (d0,d1) = a*b + (c0,c1)            ; sets CC.   d must not alias inputs (3-operand form)
#define ZKB_MADW_FIRST(d0, d1, a, b, c0, c1)                                                 \
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.cc.u32 %1, %2, %3, %5;"             \
               : "=&r"(d0), "=&r"(d1) : "r"(a), "r"(b), "r"(c0), "r"(c1))
// (d0,d1) = a*b + (c0,c1) + CC       ; sets CC
#define ZKB_MADW_CC(d0, d1, a, b, c0, c1)                                                    \
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.cc.u32 %1, %2, %3, %5;"            \
               : "=&r"(d0), "=&r"(d1) : "r"(a), "r"(b), "r"(c0), "r"(c1))
// in-place variants: (d0,d1) += a*b [+ CC]
#define ZKB_MACW_FIRST(d0, d1, a, b)                                                         \
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"             \
               : "+r"(d0), "+r"(d1) : "r"(a), "r"(b))
#define ZKB_MACW_CC(d0, d1, a, b)                                                            \
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"            \
               : "+r"(d0), "+r"(d1) : "r"(a), "r"(b))
#define ZKB_ADD_CC(d, a) asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(d) : "r"(a))
#define ZKB_ADDC_CC(d, a) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(d) : "r"(a))
#define ZKB_ADDC(d, a) asm volatile("addc.u32 %0, %0, %1;" : "+r"(d) : "r"(a))
#define ZKB_SUB_CC(d, a) asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(d) : "r"(a))
#define ZKB_SUBC_CC(d, a) asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(d) : "r"(a))
#define ZKB_SUBC(d, a) asm volatile("subc.u32 %0, %0, %1;" : "+r"(d) : "r"(a))
#else  // host emulation of the same PTX semantics (tests/emu/emu_cuda.h)
#define ZKB_MADW_FIRST(d0, d1, a, b, c0, c1) emu::madw(d0, d1, a, b, c0, c1, false)
#define ZKB_MADW_CC(d0, d1, a, b, c0, c1) emu::madw(d0, d1, a, b, c0, c1, true)
#define ZKB_MACW_FIRST(d0, d1, a, b) emu::madw(d0, d1, a, b, d0, d1, false)
#define ZKB_MACW_CC(d0, d1, a, b) emu::madw(d0, d1, a, b, d0, d1, true)
#define ZKB_ADD_CC(d, a) emu::addc(d, a, false, true)
#define ZKB_ADDC_CC(d, a) emu::addc(d, a, true, true)
#define ZKB_ADDC(d, a) emu::addc(d, a, true, false)
#define ZKB_SUB_CC(d, a) emu::subc(d, a, false, true)
#define ZKB_SUBC_CC(d, a) emu::subc(d, a, true, true)
#define ZKB_SUBC(d, a) emu::subc(d, a, true, false)
#endif

template <int L, int T>
struct Mont {
  static_assert(L % 2 == 0, "limbs per lane must be even");
  static constexpr unsigned FULL = 0xffffffffu;

  // One CIOS step.  On entry Xp holds the previous step's even-aligned array (its limb 0 already
  // handed to the lane below, the lane above's limb already added at Xp[L]) and Yp the odd-aligned
  // one; on exit the roles are swapped: Yp is the even-aligned array, Xp the odd-aligned one.
  __device__ __forceinline__ static void step(uint32_t (&Xp)[L + 2], uint32_t (&Yp)[L + 2],
                                              const uint32_t (&a)[L], const uint32_t (&n)[L],
                                              uint32_t bi, uint32_t n0inv, bool top_lane) {
    // A: the surviving high half of the cancelled pair joins the new even-aligned array
    ZKB_ADD_CC(Yp[0], Xp[1]);
    // B: new odd-aligned array, shifted down one register pair, += a_odd * bi   (carry from A)
#pragma unroll
    for (int k = 0; k < L / 2; k++)
      ZKB_MADW_CC(Xp[2 * k], Xp[2 * k + 1], a[2 * k + 1], bi, Xp[2 * k + 2], Xp[2 * k + 3]);
    Xp[L] = 0;
    ZKB_ADDC(Xp[L], 0u);
    Xp[L + 1] = 0;
    // C: new even-aligned array += a_even * bi
    ZKB_MACW_FIRST(Yp[0], Yp[1], a[0], bi);
#pragma unroll
    for (int k = 1; k < L / 2; k++) ZKB_MACW_CC(Yp[2 * k], Yp[2 * k + 1], a[2 * k], bi);
    ZKB_ADDC_CC(Yp[L], 0u);
    ZKB_ADDC(Yp[L + 1], 0u);
    // D: Montgomery quotient digit from lane 0's low limb
    uint32_t m = __shfl_sync(FULL, Yp[0] * n0inv, 0, T);
    // E: even-aligned += n_even * m   (lane 0: limb 0 becomes zero)
    ZKB_MACW_FIRST(Yp[0], Yp[1], n[0], m);
#pragma unroll
    for (int k = 1; k < L / 2; k++) ZKB_MACW_CC(Yp[2 * k], Yp[2 * k + 1], n[2 * k], m);
    ZKB_ADDC_CC(Yp[L], 0u);
    ZKB_ADDC(Yp[L + 1], 0u);
    // F: odd-aligned += n_odd * m
    ZKB_MACW_FIRST(Xp[0], Xp[1], n[1], m);
#pragma unroll
    for (int k = 1; k < L / 2; k++) ZKB_MACW_CC(Xp[2 * k], Xp[2 * k + 1], n[2 * k + 1], m);
    ZKB_ADDC_CC(Xp[L], 0u);
    ZKB_ADDC(Xp[L + 1], 0u);
    // G: the low limb of the lane above lands at limb L of this lane (weight 2^(32*(L-1)) after
    // the implicit shift of the next step)
    uint32_t recv = __shfl_down_sync(FULL, Yp[0], 1, T);
    if (top_lane) recv = 0;
    ZKB_ADD_CC(Yp[L], recv);
    ZKB_ADDC(Yp[L + 1], 0u);
  }

  // r -= (cond ? n : 0) across the T lanes of the group (borrow out of the top lane dropped).
  __device__ __forceinline__ static void cond_sub(uint32_t (&r)[L], const uint32_t (&n)[L],
                                                  bool cond, int lane_in_group, int group_shift) {
    uint32_t mask = cond ? 0xffffffffu : 0u;
    ZKB_SUB_CC(r[0], n[0] & mask);
#pragma unroll
    for (int j = 1; j < L; j++) ZKB_SUBC_CC(r[j], n[j] & mask);
    uint32_t bout = 0;
    ZKB_SUBC(bout, 0u);  // 0 - 0 - borrow = -borrow
    uint32_t orv = 0;
#pragma unroll
    for (int j = 0; j < L; j++) orv |= r[j];
    unsigned G = __ballot_sync(FULL, bout != 0), P = __ballot_sync(FULL, orv == 0);
    if (T > 1) {
      const unsigned gm = (T >= 32) ? 0xffffffffu : ((1u << T) - 1u);
      unsigned g = (G >> group_shift) & gm, p = (P >> group_shift) & gm;
      p &= ~g;
      unsigned bins = ((g + (g | p)) ^ (g ^ (g | p)));
      uint32_t bin = (bins >> lane_in_group) & 1u;
      ZKB_SUB_CC(r[0], bin);
#pragma unroll
      for (int j = 1; j < L; j++) ZKB_SUBC_CC(r[j], 0u);
    }
  }

  // r = a * b * R^-1 mod n, kept < R (one conditional subtraction on overflow).  b is read lane by
  // lane through shuffles, so passing b == a squares.
  __device__ __forceinline__ static void mul(uint32_t (&r)[L], const uint32_t (&a)[L],
                                             const uint32_t (&b)[L], const uint32_t (&n)[L],
                                             uint32_t n0inv, int lane_in_group, int group_shift) {
    uint32_t E[L + 2], O[L + 2];
#pragma unroll
    for (int j = 0; j < L + 2; j++) { E[j] = 0; O[j] = 0; }
    const bool top_lane = lane_in_group == T - 1;
#pragma unroll 1
    for (int src = 0; src < T; src++) {
#pragma unroll
      for (int jj = 0; jj < L; jj += 2) {
        uint32_t b0 = __shfl_sync(FULL, b[jj], src, T);
        uint32_t b1 = __shfl_sync(FULL, b[jj + 1], src, T);
        step(E, O, a, n, b0, n0inv, top_lane);
        step(O, E, a, n, b1, n0inv, top_lane);
      }
    }
    // merge: lane value = O + E[1] + 2^32 * (E >> 64)  -> L limbs + hi (2 limbs)
    uint32_t hi0, hi1;
    r[0] = O[0];
    ZKB_ADD_CC(r[0], E[1]);
#pragma unroll
    for (int j = 1; j < L; j++) { r[j] = O[j]; ZKB_ADDC_CC(r[j], E[j + 1]); }
    hi0 = O[L];
    ZKB_ADDC_CC(hi0, E[L + 1]);
    hi1 = O[L + 1];
    ZKB_ADDC(hi1, 0u);
    // inter-lane resolve, phase 1: add the hi limbs of the lane below
    uint32_t in0 = __shfl_up_sync(FULL, hi0, 1, T), in1 = __shfl_up_sync(FULL, hi1, 1, T);
    if (lane_in_group == 0) { in0 = 0; in1 = 0; }
    ZKB_ADD_CC(r[0], in0);
    ZKB_ADDC_CC(r[1], in1);
#pragma unroll
    for (int j = 2; j < L; j++) ZKB_ADDC_CC(r[j], 0u);
    uint32_t gout = 0;
    ZKB_ADDC(gout, 0u);
    // phase 2: single-bit carries via generate/propagate over the group's ballot bits
    uint32_t andv = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < L; j++) andv &= r[j];
    unsigned G = __ballot_sync(FULL, gout != 0), P = __ballot_sync(FULL, andv == 0xffffffffu);
    const unsigned gm = (T >= 32) ? 0xffffffffu : ((1u << T) - 1u);
    unsigned g = (G >> group_shift) & gm, p = (P >> group_shift) & gm;
    p &= ~g;
    unsigned long long s = (unsigned long long)g + (g | p);
    unsigned long long cins = s ^ (g ^ (g | p));  // bit q = carry into lane q, bit T = carry out
    uint32_t cin = (uint32_t)(cins >> lane_in_group) & 1u;
    if (T > 1) {
      ZKB_ADD_CC(r[0], cin);
#pragma unroll
      for (int j = 1; j < L; j++) ZKB_ADDC_CC(r[j], 0u);
    }
    // overflow beyond R: hi of the top lane + carry out
    uint32_t ov_top = hi0 + ((uint32_t)(cins >> T) & 1u);  // hi1 is provably zero on the top lane
    uint32_t ov = __shfl_sync(FULL, ov_top, T - 1, T);
    cond_sub(r, n, ov != 0, lane_in_group, group_shift);
  }

  // true iff a < b (group-wide, lexicographic from the top limb/lane)
  __device__ __forceinline__ static bool less_than(const uint32_t (&a)[L], const uint32_t (&b)[L],
                                                   int group_shift) {
    bool lt = false, eq = true;
#pragma unroll
    for (int j = L - 1; j >= 0; j--) {
      lt = lt || (eq && a[j] < b[j]);
      eq = eq && (a[j] == b[j]);
    }
    unsigned LT = __ballot_sync(FULL, lt), EQ = __ballot_sync(FULL, eq);
    const unsigned gm = (T >= 32) ? 0xffffffffu : ((1u << T) - 1u);
    unsigned ltg = (LT >> group_shift) & gm, neq = (~(EQ >> group_shift)) & gm;
    if (neq == 0) return false;
    int top = 31 - __clz(neq);
    return (ltg >> top) & 1u;
  }
};

// Expected EMSA-PKCS1-v1_5 limb g (little-endian limb index) for modulus byte length k and the
// SHA-256 state words h[0..7] (native words: limb i = h[7-i] for i < 8).
__device__ __forceinline__ uint32_t emsa_limb(int g, int k, const uint32_t* __restrict__ h) {
  if (g < 8) return h[7 - g];
  // bytes from the least-significant end: 32..50 = DigestInfo prefix reversed, 51 = 00,
  // 52..k-3 = FF, k-2 = 01, k-1 = 00, >= k: 00
  const uint8_t pre_rev[20] = {0x20, 0x04, 0x00, 0x05, 0x01, 0x02, 0x04, 0x03, 0x65, 0x01,
                               0x48, 0x86, 0x60, 0x09, 0x06, 0x0d, 0x30, 0x31, 0x30, 0x00};
  uint32_t v = 0;
#pragma unroll
  for (int bi = 0; bi < 4; bi++) {
    int off = 4 * g + bi;
    uint32_t byte;
    if (off < 52) byte = pre_rev[off - 32];
    else if (off < k - 2) byte = 0xff;
    else if (off == k - 2) byte = 0x01;
    else byte = 0x00;
    v |= byte << (8 * bi);
  }
  return v;
}

// One group of T lanes per work item.  GENERIC=false requires e == 65537 for every item.
template <int LIMBS, int T, bool GENERIC>
__global__ void __launch_bounds__(128)
rsa_verify_kernel(const uint32_t* __restrict__ sig_arena, const RsaItem* __restrict__ items,
                  uint32_t n_items, const uint32_t* __restrict__ keytab,
                  const uint32_t* __restrict__ digests, uint32_t* __restrict__ cand_flags) {
  constexpr int L = LIMBS / T;
  using M = Mont<L, T>;
  const int lane = threadIdx.x & 31;
  const int lig = lane % T;
  const int gshift = lane - lig;
  uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) / T;
  bool active = gid < n_items;
  if (!active) gid = n_items - 1;  // keep the whole warp in the shuffles
  const RsaItem it = items[gid];
  const uint32_t* key = keytab + (size_t)it.key_id * ZKB_KEY_STRIDE;
  uint32_t n[L], s[L], x[L];
#pragma unroll
  for (int j = 0; j < L; j++) {
    n[j] = __ldg(key + lig * L + j);
    s[j] = __ldg(sig_arena + it.sig_off + lig * L + j);
    x[j] = __ldg(key + ZKB_KEY_RR + lig * L + j);
  }
  const uint32_t n0inv = __ldg(key + ZKB_KEY_N0INV);
  const int k = (int)__ldg(key + ZKB_KEY_K);
  bool in_range = M::less_than(s, n, gshift);  // rsa: sig >= n  => Err(Verification)
  // x = s * R mod n
  M::mul(x, s, x, n, n0inv, lig, gshift);
  if (!GENERIC) {
#pragma unroll 1
    for (int i = 0; i < 16; i++) M::mul(x, x, x, n, n0inv, lig, gshift);
    M::mul(x, x, s, n, n0inv, lig, gshift);  // * s, leaves Montgomery form: s^65537 (< 2n)
  } else {
    uint64_t e = ((uint64_t)__ldg(key + ZKB_KEY_EHI) << 32) | __ldg(key + ZKB_KEY_ELO);
    uint32_t xr[L];
#pragma unroll
    for (int j = 0; j < L; j++) xr[j] = x[j];
    int top = 63 - __clzll((long long)e);
#pragma unroll 1
    for (int bit = top - 1; bit >= 0; bit--) {
      M::mul(x, x, x, n, n0inv, lig, gshift);
      // multiply unconditionally and select: groups of one warp may hold different exponents and
      // the shuffles inside mul() need every lane of the warp
      uint32_t t[L];
      M::mul(t, x, xr, n, n0inv, lig, gshift);
      if ((e >> bit) & 1) {
#pragma unroll
        for (int j = 0; j < L; j++) x[j] = t[j];
      }
    }
    uint32_t one[L];
#pragma unroll
    for (int j = 0; j < L; j++) one[j] = (lig == 0 && j == 0) ? 1u : 0u;
    M::mul(x, x, one, n, n0inv, lig, gshift);  // leave Montgomery form (<= n)
  }
  bool ge = !M::less_than(x, n, gshift);
  M::cond_sub(x, n, ge, lig, gshift);
  // compare with the expected encoded message
  const uint32_t* h = digests + (size_t)it.digest_slot * 8;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < L; j++) ok = ok && (x[j] == emsa_limb(lig * L + j, k, h));
  unsigned OK = __ballot_sync(0xffffffffu, ok);
  const unsigned gm = (T >= 32) ? 0xffffffffu : ((1u << T) - 1u);
  bool all_ok = ((OK >> gshift) & gm) == gm;
  if (active && lig == 0) {
    uint32_t f = (all_ok && in_range && k >= 62) ? ZKB_F_RSA_OK : 0u;
    atomicOr(cand_flags + it.cand, f);
  }
}

}  // namespace zkb
