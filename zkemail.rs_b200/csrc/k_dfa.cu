#include "kernels.h"
#include "dfa.cuh"
namespace zkb {
#define ZKB_DFA_DISPATCH(KERNEL, ...)                                                     \
  do {                                                                                    \
    if (elem == 2) { if (direct) KERNEL<uint16_t, true><<<grid, 128, use_smem ? smem : 0, s>>>(__VA_ARGS__);   \
                     else KERNEL<uint16_t, false><<<grid, 128, use_smem ? smem : 0, s>>>(__VA_ARGS__); }       \
    else { if (direct) KERNEL<uint32_t, true><<<grid, 128, use_smem ? smem : 0, s>>>(__VA_ARGS__);             \
           else KERNEL<uint32_t, false><<<grid, 128, use_smem ? smem : 0, s>>>(__VA_ARGS__); }                 \
  } while (0)

cudaError_t dfa_set_smem_limit(size_t bytes) {
  cudaError_t e;
#define ZKB_SET(K) if ((e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes))) return e
  ZKB_SET((dfa_scan_strided<uint16_t, true>)); ZKB_SET((dfa_scan_strided<uint16_t, false>));
  ZKB_SET((dfa_scan_strided<uint32_t, true>)); ZKB_SET((dfa_scan_strided<uint32_t, false>));
  ZKB_SET((dfa_scan_kernel<uint16_t, true>)); ZKB_SET((dfa_scan_kernel<uint16_t, false>));
  ZKB_SET((dfa_scan_kernel<uint32_t, true>)); ZKB_SET((dfa_scan_kernel<uint32_t, false>));
#undef ZKB_SET
  return cudaSuccess;
}
void launch_dfa(uint32_t elem, bool direct, const uint8_t* arena, const DfaItem* items, uint32_t n_items, const uint8_t* fwd_blob,
                uint32_t fwd_bytes, const uint8_t* rev_blob, uint32_t rev_bytes, size_t smem_limit, int qp, uint4* out,
                cudaStream_t s) {
  if (!n_items) return;
  size_t smem = (size_t)fwd_bytes + rev_bytes;
  int use_smem = smem <= smem_limit ? 1 : 0;
  unsigned grid = (n_items + 127) / 128;
  ZKB_DFA_DISPATCH(dfa_scan_kernel, arena, items, n_items, fwd_blob, fwd_bytes, rev_blob, rev_bytes, use_smem, qp, out);
}
void launch_dfa_strided(uint32_t elem, bool direct, const uint8_t* arena, const DfaItem* items, uint32_t n_emails, uint32_t which,
                        uint32_t P, uint32_t pi, const uint8_t* fwd_blob, uint32_t fwd_bytes, const uint8_t* rev_blob,
                        uint32_t rev_bytes, size_t smem_limit, int qp, uint4* out, cudaStream_t s) {
  if (!n_emails) return;
  size_t smem = (size_t)fwd_bytes + rev_bytes;
  int use_smem = smem <= smem_limit ? 1 : 0;
  unsigned grid = (n_emails + 127) / 128;
  ZKB_DFA_DISPATCH(dfa_scan_strided, arena, items, n_emails, which, P, pi, fwd_blob, fwd_bytes, rev_blob, rev_bytes, use_smem, qp, out);
}
}  // namespace zkb
