#include "kernels.h"
#include "rsa.cuh"
#include <cstdlib>
#ifndef ZKB_SQR_DEFAULT
#define ZKB_SQR_DEFAULT 8
#endif
namespace zkb {
template <bool G>
static void launch_t(int lanes, const uint32_t* sig_arena, const RsaItem* items, uint32_t n, const uint32_t* keytab,
                     const uint32_t* digests, uint32_t* cand_flags, cudaStream_t s) {
  const unsigned block = 128;
#define ZKB_RSA_CASE(TT)                                                                                  \
  case TT: {                                                                                              \
    unsigned grid = (unsigned)(((uint64_t)n * TT + block - 1) / block);                                   \
    rsa_verify_kernel<64, TT, G><<<grid, block, 0, s>>>(sig_arena, items, n, keytab, digests, cand_flags); \
    break;                                                                                                \
  }
  switch (lanes) { ZKB_RSA_CASE(2) ZKB_RSA_CASE(8) ZKB_RSA_CASE(16) default: ZKB_RSA_CASE(4) }
#undef ZKB_RSA_CASE
}
void launch_rsa64(bool generic, int lanes, const uint32_t* sig_arena, const RsaItem* items, uint32_t n,
                  const uint32_t* keytab, const uint32_t* digests, uint32_t* cand_flags, cudaStream_t s, bool sqr) {
  if (!n) return;
  if (sqr && !generic && lanes == 4) {
    // e = 65537, four lanes per signature: the variant with the dedicated squaring (Mont::sqr), 64-thread CTAs
    const unsigned block = 64;
    const unsigned grid = (unsigned)(((uint64_t)n * 4 + block - 1) / block);
    // 27 KB of shared memory per CTA: ask for the largest carve-out, otherwise shared memory (not registers) caps the
    // kernel at 6 CTAs per SM (ncu: launch__occupancy_limit_shared_mem).
#define ZKB_SQR_CASE(V)                                                                                             \
  case V: {                                                                                                         \
    /* a per-device attribute: set at every launch (engines of several devices share this code) */                  \
    cudaFuncSetAttribute(rsa_verify_kernel<64, 4, false, V>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);  \
    rsa_verify_kernel<64, 4, false, V><<<grid, block, 0, s>>>(sig_arena, items, n, keytab, digests, cand_flags);    \
    break;                                                                                                          \
  }
#ifdef ZKB_SQR_EXPERIMENTS
    // A/B builds only (make EXTRA=-DZKB_SQR_EXPERIMENTS): ZKB_SQR_VARIANT picks one of the code-shape variants of rsa.cuh.
    // The product library instantiates the measured best one and reads no environment variables.
    static int variant = -1;
    if (variant < 0) { const char* v = getenv("ZKB_SQR_VARIANT"); variant = v ? atoi(v) : ZKB_SQR_DEFAULT; }
    switch (variant) { ZKB_SQR_CASE(4) ZKB_SQR_CASE(16) ZKB_SQR_CASE(104) ZKB_SQR_CASE(108) ZKB_SQR_CASE(1008) ZKB_SQR_CASE(2008) default: ZKB_SQR_CASE(8) }
#else
    switch (ZKB_SQR_DEFAULT) { default: ZKB_SQR_CASE(ZKB_SQR_DEFAULT) }
#endif
#undef ZKB_SQR_CASE
    return;
  }
  if (generic) launch_t<true>(lanes, sig_arena, items, n, keytab, digests, cand_flags, s);
  else launch_t<false>(lanes, sig_arena, items, n, keytab, digests, cand_flags, s);
}
}  // namespace zkb
