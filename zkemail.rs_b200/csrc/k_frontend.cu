#include "kernels.h"
#include "frontend.cuh"
namespace zkb {
void launch_frontend(const uint8_t* span, const FeIn* in, uint32_t n, uint8_t* arena, const uint64_t* msg_off, uint32_t* msg_len,
                     uint32_t* sig_arena, uint32_t* cand_bh, CanonItem* canon, FeOut* out, cudaStream_t s) {
  if (n) frontend_kernel<<<(n + 127) / 128, 128, 0, s>>>(span, in, n, arena, msg_off, msg_len, sig_arena, cand_bh, canon, out);
}
}  // namespace zkb
