#include "../../include/zkemail_b200.h"
#include "kernels.h"
#include "assemble.cuh"
namespace zkb {
static_assert(offsetof(zkb_result, parts) == ZKB_REC_HEAD, "record head must be the head of zkb_result");
static_assert(sizeof(((zkb_result*)0)->parts[0]) == 16, "part entry is one uint4");
void launch_assemble(const FeIn* in, const FeOut* fo, uint32_t n, const uint32_t* cand_flags, const uint32_t* digests, const uint4* dfa_out,
                     uint32_t P, uint32_t body_mask, bool have_regex, uint32_t* recs, uint32_t rec_words, cudaStream_t s) {
  if (!n) return;
  assemble_kernel<<<(n + 127) / 128, 128, 0, s>>>(in, fo, n, cand_flags, digests, dfa_out, P, body_mask, have_regex ? 1u : 0u, recs, rec_words);
}
}  // namespace zkb
