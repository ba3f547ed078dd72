#include "kernels.h"
#include "frontend.cuh"
namespace zkb {
void launch_frontend(const uint8_t* span, const FeIn* in, uint32_t n, uint8_t* arena, const uint64_t* msg_off, uint32_t* msg_len,
                     uint32_t* sig_arena, uint32_t* cand_bh, CanonItem* canon, FeOut* out, bool allow_skip, long long now, cudaStream_t s) {
  if (!n) return;
  // lane = message: chunks of the e2e pipeline hold ~64 K messages, which 128-thread CTAs spread unevenly (3 or 4
  // CTAs per SM); small CTAs balance them (an SM holds 32 CTAs, so only while that does not cap the occupancy)
  const unsigned block = n <= 148u * 1024u ? 32u : n <= 148u * 2048u ? 64u : 128u;
  frontend_kernel<<<(n + block - 1) / block, block, 0, s>>>(span, in, n, arena, msg_off, msg_len, sig_arena, cand_bh, canon, out, allow_skip ? 1 : 0, now);
}
}  // namespace zkb
