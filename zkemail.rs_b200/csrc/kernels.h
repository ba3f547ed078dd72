// kernels.h — host-callable launchers of the sm_100a kernels (one translation unit per kernel
// family so that the library builds in parallel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace zkb {

void launch_sha256(const uint8_t* arena, const uint64_t* msg_off, const uint32_t* msg_len, const uint32_t* order,
                   uint32_t n, uint32_t* digests, cudaStream_t s);
void launch_bh_check(const uint32_t* digests, const uint32_t* body_slot, const uint32_t* bh_words, uint32_t n_cand,
                     uint32_t* cand_flags, cudaStream_t s);
// limbs in {32,64,128}; lanes = threads cooperating on one signature (clamped to a compiled value)
void launch_rsa(int limbs, bool generic, int lanes, const uint32_t* sig_arena, const RsaItem* items, uint32_t n,
                const uint32_t* keytab, const uint32_t* digests, uint32_t* cand_flags, cudaStream_t s);
void launch_rsa32(bool generic, int lanes, const uint32_t* sig_arena, const RsaItem* items, uint32_t n,
                  const uint32_t* keytab, const uint32_t* digests, uint32_t* cand_flags, cudaStream_t s);
// sqr: four lanes per signature and e = 65537 use the kernel variant with the dedicated Montgomery squaring
void launch_rsa64(bool generic, int lanes, const uint32_t* sig_arena, const RsaItem* items, uint32_t n,
                  const uint32_t* keytab, const uint32_t* digests, uint32_t* cand_flags, cudaStream_t s, bool sqr = false);
void launch_rsa128(bool generic, int lanes, const uint32_t* sig_arena, const RsaItem* items, uint32_t n,
                   const uint32_t* keytab, const uint32_t* digests, uint32_t* cand_flags, cudaStream_t s);
cudaError_t dfa_set_smem_limit(size_t bytes);
void launch_dfa(uint32_t elem, bool direct, const uint8_t* arena, const DfaItem* items, uint32_t n_items, const uint32_t* msg_len, const uint8_t* fwd_blob,
                uint32_t fwd_bytes, const uint8_t* rev_blob, uint32_t rev_bytes, size_t smem_limit, int qp, uint4* out,
                cudaStream_t s);
void launch_dfa_strided(uint32_t elem, bool direct, const uint8_t* arena, const DfaItem* items, uint32_t n_emails, const uint32_t* msg_len, uint32_t which,
                        uint32_t P, uint32_t pi, const uint8_t* fwd_blob, uint32_t fwd_bytes, const uint8_t* rev_blob,
                        uint32_t rev_bytes, size_t smem_limit, int qp, uint4* out, cudaStream_t s);
// order / msg_canon (optional): walk the messages in the SHA order (descending length) and canonicalise those that are
// bodies, so that the 32 bodies of a warp have similar lengths; without them items are taken in index order
void launch_canon_body(const uint8_t* span, const CanonItem* items, uint32_t n, uint8_t* arena, const uint64_t* msg_off,
                       uint32_t* msg_len, const uint32_t* order, const uint32_t* msg_canon, uint32_t n_msgs, cudaStream_t s);
void launch_frontend(const uint8_t* span, const FeIn* in, uint32_t n, uint8_t* arena, const uint64_t* msg_off, uint32_t* msg_len,
                     uint32_t* sig_arena, uint32_t* cand_bh, CanonItem* canon, FeOut* out, bool allow_skip, long long now, cudaStream_t s);
// per-email result records (assemble.cuh); rec_words = record stride in 32-bit words
void launch_assemble(const FeIn* in, const FeOut* fo, uint32_t n, const uint32_t* cand_flags, const uint32_t* digests, const uint4* dfa_out,
                     uint32_t P, uint32_t body_mask, bool have_regex, uint32_t* recs, uint32_t rec_words, cudaStream_t s);
void launch_int_peak(int kind, unsigned grid, unsigned block, uint32_t* out, uint32_t seed, int iters, cudaStream_t s);

}  // namespace zkb
