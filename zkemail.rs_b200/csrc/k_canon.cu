#include "kernels.h"
#include "canon.cuh"
namespace zkb {
void launch_canon_body(const uint8_t* span, const CanonItem* items, uint32_t n, uint8_t* arena, const uint64_t* msg_off,
                       uint32_t* msg_len, const uint32_t* order, const uint32_t* msg_canon, uint32_t n_msgs, cudaStream_t s) {
  if (!n) return;
  // lane = message: chunks of the e2e pipeline hold ~64 K messages, which 128-thread CTAs spread unevenly (3 or 4
  // CTAs per SM); small CTAs balance them (an SM holds 32 CTAs, so only while that does not cap the occupancy)
  // staged form: lane = body, bytes move through shared memory in coalesced 128-byte segments (canon.cuh)
  const unsigned per_cta = CANON_WARPS * 32;
  const uint32_t lanes = (order && msg_canon) ? n_msgs : n;
  canon_body_staged_kernel<<<(lanes + per_cta - 1) / per_cta, per_cta, 0, s>>>(span, items, n, arena, msg_off, msg_len, order, msg_canon, lanes);
}
}  // namespace zkb
