#include "kernels.h"
#include "frontend_warp.cuh"
namespace zkb {
void launch_frontend(const uint8_t* span, const FeIn* in, uint32_t n, uint8_t* arena, const uint64_t* msg_off, uint32_t* msg_len,
                     uint32_t* sig_arena, uint32_t* cand_bh, CanonItem* canon, FeOut* out, bool allow_skip, long long now, cudaStream_t s) {
  if (!n) return;
  // one warp per message, FE_WARPS messages per CTA (frontend_warp.cuh); the scalar lane-per-message form
  // (frontend.cuh: frontend_kernel) stays in the tree as the twin the emulated tests compare against
  frontend_warp_kernel<<<(n + FE_WARPS - 1) / FE_WARPS, FE_WARPS * 32, 0, s>>>(span, in, n, arena, msg_off, msg_len, sig_arena, cand_bh, canon, out,
                                                                            allow_skip ? 1 : 0, now);
}
}  // namespace zkb
