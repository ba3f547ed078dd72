#include "kernels.h"
#include "dfa.cuh"
namespace zkb {
cudaError_t dfa_set_smem_limit(size_t bytes) {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(dfa_scan_strided<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes))) return e;
  if ((e = cudaFuncSetAttribute(dfa_scan_strided<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes))) return e;
  if ((e = cudaFuncSetAttribute(dfa_scan_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes))) return e;
  return cudaFuncSetAttribute(dfa_scan_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
void launch_dfa(uint32_t elem, const uint8_t* arena, const DfaItem* items, uint32_t n_items, const uint8_t* fwd_blob,
                uint32_t fwd_bytes, const uint8_t* rev_blob, uint32_t rev_bytes, size_t smem_limit, int qp, uint4* out,
                cudaStream_t s) {
  if (!n_items) return;
  size_t smem = (size_t)fwd_bytes + rev_bytes;
  int use_smem = smem <= smem_limit ? 1 : 0;
  unsigned grid = (n_items + 127) / 128;
  if (elem == 2) dfa_scan_kernel<uint16_t><<<grid, 128, use_smem ? smem : 0, s>>>(arena, items, n_items, fwd_blob, fwd_bytes, rev_blob, rev_bytes, use_smem, qp, out);
  else dfa_scan_kernel<uint32_t><<<grid, 128, use_smem ? smem : 0, s>>>(arena, items, n_items, fwd_blob, fwd_bytes, rev_blob, rev_bytes, use_smem, qp, out);
}
void launch_dfa_strided(uint32_t elem, const uint8_t* arena, const DfaItem* items, uint32_t n_emails, uint32_t which,
                        uint32_t P, uint32_t pi, const uint8_t* fwd_blob, uint32_t fwd_bytes, const uint8_t* rev_blob,
                        uint32_t rev_bytes, size_t smem_limit, int qp, uint4* out, cudaStream_t s) {
  if (!n_emails) return;
  size_t smem = (size_t)fwd_bytes + rev_bytes;
  int use_smem = smem <= smem_limit ? 1 : 0;
  unsigned grid = (n_emails + 127) / 128;
  if (elem == 2) dfa_scan_strided<uint16_t><<<grid, 128, use_smem ? smem : 0, s>>>(arena, items, n_emails, which, P, pi, fwd_blob, fwd_bytes, rev_blob, rev_bytes, use_smem, qp, out);
  else dfa_scan_strided<uint32_t><<<grid, 128, use_smem ? smem : 0, s>>>(arena, items, n_emails, which, P, pi, fwd_blob, fwd_bytes, rev_blob, rev_bytes, use_smem, qp, out);
}
}  // namespace zkb
