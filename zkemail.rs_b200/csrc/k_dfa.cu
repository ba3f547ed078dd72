#include "kernels.h"
#include "dfa.cuh"
namespace zkb {
#define ZKB_DFA_GO(TT, D, S, KERNEL, ...) KERNEL<TT, D, S><<<grid, 128, S ? smem : 0, s>>>(__VA_ARGS__)
#define ZKB_DFA_DISPATCH(KERNEL, ...)                                                              \
  do {                                                                                             \
    const int sel = (elem == 2 ? 0 : 4) | (direct ? 2 : 0) | (use_smem ? 1 : 0);                   \
    switch (sel) {                                                                                 \
      case 0: ZKB_DFA_GO(uint16_t, false, false, KERNEL, __VA_ARGS__); break;                      \
      case 1: ZKB_DFA_GO(uint16_t, false, true, KERNEL, __VA_ARGS__); break;                       \
      case 2: ZKB_DFA_GO(uint16_t, true, false, KERNEL, __VA_ARGS__); break;                       \
      case 3: ZKB_DFA_GO(uint16_t, true, true, KERNEL, __VA_ARGS__); break;                        \
      case 4: ZKB_DFA_GO(uint32_t, false, false, KERNEL, __VA_ARGS__); break;                      \
      case 5: ZKB_DFA_GO(uint32_t, false, true, KERNEL, __VA_ARGS__); break;                       \
      case 6: ZKB_DFA_GO(uint32_t, true, false, KERNEL, __VA_ARGS__); break;                       \
      default: ZKB_DFA_GO(uint32_t, true, true, KERNEL, __VA_ARGS__); break;                       \
    }                                                                                              \
  } while (0)

cudaError_t dfa_set_smem_limit(size_t bytes) {
  cudaError_t e;
#define ZKB_SET(K) if ((e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes))) return e
  ZKB_SET((dfa_scan_strided<uint16_t, true, true>)); ZKB_SET((dfa_scan_strided<uint16_t, false, true>));
  ZKB_SET((dfa_scan_strided<uint32_t, true, true>)); ZKB_SET((dfa_scan_strided<uint32_t, false, true>));
  ZKB_SET((dfa_scan_kernel<uint16_t, true, true>)); ZKB_SET((dfa_scan_kernel<uint16_t, false, true>));
  ZKB_SET((dfa_scan_kernel<uint32_t, true, true>)); ZKB_SET((dfa_scan_kernel<uint32_t, false, true>));
#undef ZKB_SET
  return cudaSuccess;
}
void launch_dfa(uint32_t elem, bool direct, const uint8_t* arena, const DfaItem* items, uint32_t n_items, const uint32_t* msg_len, const uint8_t* fwd_blob,
                uint32_t fwd_bytes, const uint8_t* rev_blob, uint32_t rev_bytes, size_t smem_limit, int qp, uint4* out,
                cudaStream_t s) {
  if (!n_items) return;
  size_t smem = (size_t)fwd_bytes + rev_bytes;
  const bool use_smem = smem <= smem_limit;
  unsigned grid = (n_items + 127) / 128;
  ZKB_DFA_DISPATCH(dfa_scan_kernel, arena, items, n_items, msg_len, fwd_blob, fwd_bytes, rev_blob, rev_bytes, qp, out);
}
void launch_dfa_strided(uint32_t elem, bool direct, const uint8_t* arena, const DfaItem* items, uint32_t n_emails, const uint32_t* msg_len, uint32_t which,
                        uint32_t P, uint32_t pi, const uint8_t* fwd_blob, uint32_t fwd_bytes, const uint8_t* rev_blob,
                        uint32_t rev_bytes, size_t smem_limit, int qp, uint4* out, cudaStream_t s) {
  if (!n_emails) return;
  size_t smem = (size_t)fwd_bytes + rev_bytes;
  const bool use_smem = smem <= smem_limit;
  unsigned grid = (n_emails + 127) / 128;
  ZKB_DFA_DISPATCH(dfa_scan_strided, arena, items, n_emails, msg_len, which, P, pi, fwd_blob, fwd_bytes, rev_blob, rev_bytes, qp, out);
}
}  // namespace zkb
