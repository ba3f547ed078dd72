#include "kernels.h"
#include "canon.cuh"
namespace zkb {
void launch_canon_body(const uint8_t* span, const CanonItem* items, uint32_t n, uint8_t* arena, const uint64_t* msg_off,
                       uint32_t* msg_len, cudaStream_t s) {
  if (n) canon_body_kernel<<<(n + 127) / 128, 128, 0, s>>>(span, items, n, arena, msg_off, msg_len);
}
}  // namespace zkb
