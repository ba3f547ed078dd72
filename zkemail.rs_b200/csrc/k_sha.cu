#include "kernels.h"
#include "sha256.cuh"
namespace zkb {
void launch_sha256(const uint8_t* arena, const uint64_t* msg_off, const uint32_t* msg_len, const uint32_t* order,
                   uint32_t n, uint32_t* digests, cudaStream_t s) {
  if (n) sha256_batch_kernel<<<(n + 127) / 128, 128, 0, s>>>(arena, msg_off, msg_len, order, n, digests, 1u);
}
void launch_bh_check(const uint32_t* digests, const uint32_t* body_slot, const uint32_t* bh_words, uint32_t n_cand,
                     uint32_t* cand_flags, cudaStream_t s) {
  if (n_cand) bh_check_kernel<<<(n_cand + 255) / 256, 256, 0, s>>>(digests, body_slot, bh_words, n_cand, cand_flags);
}
}  // namespace zkb
