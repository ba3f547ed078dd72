#include "kernels.h"
#include "sha256.cuh"
#include <cstdlib>
namespace zkb {
void launch_sha256(const uint8_t* arena, const uint64_t* msg_off, const uint32_t* msg_len, const uint32_t* order,
                   uint32_t n, uint32_t* digests, cudaStream_t s) {
  if (!n) return;
  // lane = message.  Batches with few (large) messages use smaller CTAs so that the CTAs spread evenly over the
  // 148 SMs (100 k messages are 782 CTAs of 128 threads: 5 or 6 per SM, a 12 % imbalance; 3125 CTAs of 32 threads: 4 %);
  // an SM holds at most 32 CTAs, so small CTAs are only used while they do not cap the occupancy.
  const unsigned block = n <= 148u * 1024u ? 32u : n <= 148u * 2048u ? 64u : 128u;
  const unsigned grid = (n + block - 1) / block;
  const bool few = n <= 148u * 2048u;
#define ZKB_SHA_CASE(R)                                                                                         \
  case R:                                                                                                       \
    if (few) sha256_batch_kernel<true, R><<<grid, block, 0, s>>>(arena, msg_off, msg_len, order, n, digests, 1u); \
    else sha256_batch_kernel<false, R><<<grid, block, 0, s>>>(arena, msg_off, msg_len, order, n, digests, 1u);   \
    break;
#ifdef ZKB_SHA_EXPERIMENTS
  // A/B builds only (make EXTRA=-DZKB_SHA_EXPERIMENTS): ZKB_SHA_ROT picks how many rotate families go to the FMA pipe
  static int rot = -1;
  if (rot < 0) { const char* v = getenv("ZKB_SHA_ROT"); rot = v ? atoi(v) : ZKB_SHA_ROT_DEFAULT; }
  switch (rot) { ZKB_SHA_CASE(0) ZKB_SHA_CASE(1) ZKB_SHA_CASE(2) ZKB_SHA_CASE(3) }
#else
  switch (ZKB_SHA_ROT_DEFAULT) { ZKB_SHA_CASE(ZKB_SHA_ROT_DEFAULT) }
#endif
#undef ZKB_SHA_CASE
}
void launch_bh_check(const uint32_t* digests, const uint32_t* body_slot, const uint32_t* bh_words, uint32_t n_cand,
                     uint32_t* cand_flags, cudaStream_t s) {
  if (n_cand) bh_check_kernel<<<(n_cand + 255) / 256, 256, 0, s>>>(digests, body_slot, bh_words, n_cand, cand_flags);
}
}  // namespace zkb
