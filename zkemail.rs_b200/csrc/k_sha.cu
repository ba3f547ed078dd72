#include "kernels.h"
#include "sha256.cuh"
namespace zkb {
void launch_sha256(const uint8_t* arena, const uint64_t* msg_off, const uint32_t* msg_len, const uint32_t* order,
                   uint32_t n, uint32_t* digests, cudaStream_t s) {
  if (!n) return;
  // lane = message.  Batches with few (large) messages use smaller CTAs so that the CTAs spread evenly over the
  // 148 SMs (100 k messages are 782 CTAs of 128 threads: 5 or 6 per SM, a 12 % imbalance; 3125 CTAs of 32 threads: 4 %);
  // an SM holds at most 32 CTAs, so small CTAs are only used while they do not cap the occupancy.
  const unsigned block = n <= 148u * 1024u ? 32u : n <= 148u * 2048u ? 64u : 128u;
  if (n <= 148u * 2048u) sha256_batch_kernel<true><<<(n + block - 1) / block, block, 0, s>>>(arena, msg_off, msg_len, order, n, digests, 1u);
  else sha256_batch_kernel<false><<<(n + block - 1) / block, block, 0, s>>>(arena, msg_off, msg_len, order, n, digests, 1u);
}
void launch_bh_check(const uint32_t* digests, const uint32_t* body_slot, const uint32_t* bh_words, uint32_t n_cand,
                     uint32_t* cand_flags, cudaStream_t s) {
  if (n_cand) bh_check_kernel<<<(n_cand + 255) / 256, 256, 0, s>>>(digests, body_slot, bh_words, n_cand, cand_flags);
}
}  // namespace zkb
