// canon.cuh — device-side DKIM body canonicalisation for sm_100a (kernel K0).
//
// Replaces, on the device, cfdkim's body canonicalisation as reached from
// cfdkim::verify_email_with_key / canonicalize_signed_email (core/src/email.rs:31-33,
// core/src/circuits.rs:34-35; SURVEY.md A.2 "Relaxed body" / "Simple body"):
//   relaxed: TAB->SP, collapse SP runs, delete the SP of every " \r\n", strip trailing empty lines
//            (while the body ends in CRLF CRLF), append CRLF if non-empty and not CRLF-terminated;
//   simple : empty -> CRLF, else strip trailing empty lines;   then the optional l= truncation.
// Used when the caller's raw messages live in registered (pinned) host memory: the raw bytes are
// DMA'd to HBM as they are (no host copy, no host canonicalisation) and this kernel writes the
// canonical body into the arena slot the SHA-256 and DFA kernels read, plus its length.
//
// Mapping: lane = body.  The relaxed transform is a forward transducer with one byte of look-ahead
// (a pending SP is dropped iff the next two bytes are CR LF), so a lane streams its body with
// aligned 16-byte loads and writes the output through a 64-bit shift register (one 8-byte store per
// 8 output bytes).  A 16-byte block that needs no editing — no TAB, no SP followed by SP or CR (the
// byte before the block included); found with a handful of SWAR byte-mask operations — is passed through as two 8-byte
// emits (~2.5 instructions per byte); only blocks that contain something to canonicalise, and the two
// boundary blocks, take the byte loop (~25 instructions per byte, branch-free inside).  Simple
// canonicalisation is the pass-through for every full block.
#pragma once
#include "common.cuh"

namespace zkb {

// 0x80 in every byte of x that equals c (exact, no borrow artefacts)
__device__ __forceinline__ uint32_t canon_eq_mask(uint32_t x, uint32_t c) {
  const uint32_t t = x ^ (c * 0x01010101u);
  return ~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t) & 0x80808080u;
}

struct CanonOut {
  uint8_t* out;
  uint64_t acc;
  uint32_t o;    // bytes emitted
  __device__ __forceinline__ void emit(uint32_t c, bool on) {
    if (on) {
      acc |= (uint64_t)c << ((o & 7u) * 8u);
      o++;
      if ((o & 7u) == 0) { *reinterpret_cast<uint64_t*>(out + o - 8) = acc; acc = 0; }
    }
  }
  // eight bytes at once (the pass-through path)
  __device__ __forceinline__ void emit8(uint64_t q) {
    const uint32_t k = (o & 7u) * 8u;
    if (k == 0) { *reinterpret_cast<uint64_t*>(out + o) = q; }
    else {
      *reinterpret_cast<uint64_t*>(out + (o & ~7u)) = acc | (q << k);
      acc = q >> (64u - k);
    }
    o += 8;
  }
  __device__ __forceinline__ void flush() {
    for (uint32_t k = 0; k < (o & 7u); k++) out[(o & ~7u) + k] = (uint8_t)(acc >> (8 * k));
    acc = 0;
  }
};

__global__ void __launch_bounds__(128)
canon_body_kernel(const uint8_t* __restrict__ span, const CanonItem* __restrict__ items, uint32_t n_items,
                  uint8_t* __restrict__ arena, const uint64_t* __restrict__ msg_off, uint32_t* __restrict__ msg_len) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_items) return;
  const CanonItem it = items[idx];
  const uint8_t* in = span + it.raw_off;
  const uint32_t n = it.raw_len;
  CanonOut w;
  w.out = arena + msg_off[it.msg];
  w.acc = 0; w.o = 0;
  const bool relaxed = (it.flags & 1u) != 0;
  // aligned 16-byte blocks covering [in, in+n); bytes outside the body are predicated off
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(in) & ~(uintptr_t)15;
  const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(in) - a0);
  const uint32_t total = lead + n;
  bool pending_sp = false;   // a WSP run has been seen and not yet emitted
  int prev = -1;             // byte waiting for its look-ahead
  for (uint32_t base = 0; base < total; base += 16) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(a0 + base));
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
    // pass-through test: a full block behind a pending non-WSP byte, with nothing to edit inside
    bool pass = base >= lead && base + 16 <= total && prev >= 0;
    if (pass && relaxed) {
      pass = !pending_sp && prev != '\t';
      uint32_t sp[4], cr[4], tab = 0;
#pragma unroll
      for (int q = 0; q < 4; q++) { sp[q] = canon_eq_mask(w4[q], ' '); cr[q] = canon_eq_mask(w4[q], '\r'); tab |= canon_eq_mask(w4[q], '\t'); }
      uint32_t bad = tab;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        // flags of byte i + 1 aligned at byte i (the successor of byte 15 is judged with the next block: that byte
        // becomes `prev`, and a WSP `prev` sends the next block through the byte loop)
        const uint32_t nx = ((sp[q] | cr[q]) >> 8) | (q < 3 ? (sp[q + 1] | cr[q + 1]) << 24 : 0u);
        bad |= sp[q] & nx;
      }
      // a SP waiting as `prev` is passed through like any other byte unless byte 0 continues the run or starts a CRLF
      if (prev == ' ') bad |= (sp[0] | cr[0]) & 0x80u;
      pass = pass && bad == 0;
    }
    if (pass) {
      const uint32_t p = (uint32_t)prev;
      const uint32_t o0 = p | (v.x << 8), o1 = __funnelshift_l(v.x, v.y, 8), o2 = __funnelshift_l(v.y, v.z, 8), o3 = __funnelshift_l(v.z, v.w, 8);
      w.emit8((uint64_t)o0 | ((uint64_t)o1 << 32));
      w.emit8((uint64_t)o2 | ((uint64_t)o3 << 32));
      prev = (int)(v.w >> 24);
      continue;
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const uint32_t pos = base + (uint32_t)k;
      const bool valid = pos >= lead && pos < total;
      const uint32_t cur = (w4[k >> 2] >> ((k & 3) * 8)) & 0xffu;
      // process `prev` now that its successor `cur` is known
      const bool have = valid && prev >= 0;
      const uint32_t c = (uint32_t)prev;
      if (relaxed) {
        const bool wsp = c == ' ' || c == '\t';
        const bool crlf = c == '\r' && cur == '\n';
        const bool sp_out = have && !wsp && pending_sp && !crlf;
        w.emit(' ', sp_out);
        w.emit(c, have && !wsp);
        pending_sp = have ? wsp : pending_sp;
      } else {
        w.emit(c, have);
      }
      prev = valid ? (int)cur : prev;
    }
  }
  // the last byte has no successor
  if (prev >= 0) {
    const uint32_t c = (uint32_t)prev;
    if (relaxed) {
      const bool wsp = c == ' ' || c == '\t';
      w.emit(' ', !wsp && pending_sp);
      w.emit(c, !wsp);
      w.emit(' ', wsp);  // a trailing WSP run is kept as one SP (cfdkim quirk: "abc " -> "abc \r\n")
    } else {
      w.emit(c, true);
    }
  }
  w.flush();
  uint32_t o = w.o;
  uint8_t* out = w.out;
  if (relaxed) {
    while (o >= 4 && out[o - 1] == '\n' && out[o - 2] == '\r' && out[o - 3] == '\n' && out[o - 4] == '\r') o -= 2;
    if (o > 0 && !(o >= 2 && out[o - 2] == '\r' && out[o - 1] == '\n')) { out[o] = '\r'; out[o + 1] = '\n'; o += 2; }
  } else {
    if (n == 0) { out[0] = '\r'; out[1] = '\n'; o = 2; }
    else while (o >= 4 && out[o - 1] == '\n' && out[o - 2] == '\r' && out[o - 3] == '\n' && out[o - 4] == '\r') o -= 2;
  }
  if ((it.flags & 2u) && it.l < o) o = it.l;  // l=: Vec::truncate of the canonical body
  msg_len[it.msg] = o;
}

// ---------------------------------------------------------------------------------------------------------------------
// Staged form (the one the engine launches): the same lane-per-body transducer, but the bytes travel between HBM and the
// lanes through shared memory so that every global access is a coalesced 128-byte segment.  The lane-per-body kernel
// above had each lane stream its own body (a warp-wide load touched 32 different cache lines for 16 bytes each, 8-byte
// stores likewise): ncu showed it latency-bound at 20 % of the HBM rate with 1.7x read and 1.5x write amplification.
// Per round the warp loads, for each of its 32 bodies, the next 128 bytes (eight lanes x 16 bytes per body, four bodies
// per load instruction) into the body's shared-memory row; every lane runs the transducer over its row and emits into
// its output row; full 16-byte chunks of the output rows are written back the same cooperative way.
#define CANON_SEG 128u
#define CANON_IN_PITCH 144u    // bytes per input row (128 + 16: rows start 4 banks apart)
#define CANON_OUT_PITCH 176u   // output row: <= 15 carried + 129 emitted + an 8-byte store slot
#define CANON_WARPS 4

struct CanonOutS {             // CanonOut writing into the lane's shared-memory row
  uint8_t* row;                // row base (16-byte aligned)
  uint64_t acc;
  uint32_t o;                  // bytes emitted in total
  uint32_t flushed;            // bytes already written to global memory (multiple of 16)
  __device__ __forceinline__ void emit(uint32_t c, bool on) {
    if (on) {
      acc |= (uint64_t)c << ((o & 7u) * 8u);
      o++;
      if ((o & 7u) == 0) { *reinterpret_cast<uint64_t*>(row + (o - flushed) - 8) = acc; acc = 0; }
    }
  }
  __device__ __forceinline__ void emit8(uint64_t q) {
    const uint32_t k = (o & 7u) * 8u;
    if (k == 0) { *reinterpret_cast<uint64_t*>(row + (o - flushed)) = q; }
    else {
      *reinterpret_cast<uint64_t*>(row + ((o - flushed) & ~7u)) = acc | (q << k);
      acc = q >> (64u - k);
    }
    o += 8;
  }
  // the partial word joins the row (used before a flush so that the row holds every emitted byte)
  __device__ __forceinline__ void park() { if (o & 7u) *reinterpret_cast<uint64_t*>(row + ((o - flushed) & ~7u)) = acc; }
};

#ifdef ZKB_HOST_EMU
#define ZKB_CANON_SMEM(name) static uint8_t name[CANON_WARPS * 32 * (CANON_IN_PITCH + CANON_OUT_PITCH)] __attribute__((aligned(16)))
#else
#define ZKB_CANON_SMEM(name) __shared__ __align__(16) uint8_t name[CANON_WARPS * 32 * (CANON_IN_PITCH + CANON_OUT_PITCH)]
#endif

__global__ void __launch_bounds__(CANON_WARPS * 32)
canon_body_staged_kernel(const uint8_t* __restrict__ span, const CanonItem* __restrict__ items, uint32_t n_items,
                         uint8_t* __restrict__ arena, const uint64_t* __restrict__ msg_off, uint32_t* __restrict__ msg_len,
                         const uint32_t* __restrict__ order, const uint32_t* __restrict__ msg_canon, uint32_t n_lanes) {
  ZKB_CANON_SMEM(smem);
  const unsigned FULL = 0xffffffffu;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* in_rows = smem + warp * 32 * (CANON_IN_PITCH + CANON_OUT_PITCH);
  uint8_t* out_rows = in_rows + 32 * CANON_IN_PITCH;
  const uint32_t idx = (blockIdx.x * CANON_WARPS + warp) * 32 + lane;
  if ((blockIdx.x * CANON_WARPS + warp) * 32 >= n_lanes) return;   // whole warp out of range
  CanonItem it;
  it.raw_off = 0; it.raw_len = 0; it.msg = 0; it.flags = 0; it.l = 0;
  // lane -> item: in index order, or (order + msg_canon given) the idx-th message of the SHA order when it is a body
  uint32_t item = idx < n_lanes ? idx : 0xFFFFFFFFu;
  if (order && msg_canon && idx < n_lanes) item = msg_canon[order[idx]];
  const bool live = item < n_items;
  if (__ballot_sync(0xffffffffu, live) == 0) return;               // a warp without bodies (headers, domains, keys)
  if (live) it = items[item];
  const uint8_t* in = span + it.raw_off;
  const uint32_t n = live ? it.raw_len : 0u;
  const bool relaxed = (it.flags & 1u) != 0;
  uint8_t* gout = live ? arena + msg_off[it.msg] : arena;
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(in) & ~(uintptr_t)15;
  const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(in) - a0);
  const uint32_t total = n ? lead + n : 0u;
  uint32_t max_total = total;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { const uint32_t v = __shfl_xor_sync(FULL, max_total, d); max_total = v > max_total ? v : max_total; }
  CanonOutS w;
  w.row = out_rows + lane * CANON_OUT_PITCH;
  w.acc = 0; w.o = 0; w.flushed = 0;
  const uint8_t* my_in = in_rows + lane * CANON_IN_PITCH;
  bool pending_sp = false;
  int prev = -1;
  const uint32_t a0_lo = (uint32_t)a0, a0_hi = (uint32_t)((uint64_t)a0 >> 32);
  const uint32_t go_lo = (uint32_t)reinterpret_cast<uintptr_t>(gout), go_hi = (uint32_t)((uint64_t)reinterpret_cast<uintptr_t>(gout) >> 32);
  for (uint32_t seg = 0; seg < max_total; seg += CANON_SEG) {
    // ---- cooperative load: 4 bodies per instruction, 8 lanes x 16 bytes each
#pragma unroll
    for (int g = 0; g < 8; g++) {
      const uint32_t body = 4u * g + (lane >> 3), chunk = lane & 7u;
      const uint32_t blo = __shfl_sync(FULL, a0_lo, body), bhi = __shfl_sync(FULL, a0_hi, body), btot = __shfl_sync(FULL, total, body);
      const uint32_t off = seg + 16u * chunk;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (off < btot) v = __ldg(reinterpret_cast<const uint4*>((((uint64_t)bhi << 32) | blo) + off));
      *reinterpret_cast<uint4*>(in_rows + body * CANON_IN_PITCH + 16u * chunk) = v;
    }
    __syncwarp();
    // ---- the transducer over this lane's 8 blocks (same logic as canon_body_kernel)
#pragma unroll 1
    for (uint32_t blk = 0; blk < CANON_SEG / 16u; blk++) {
      const uint32_t base = seg + 16u * blk;
      if (base >= total) break;
      const uint4 v = *reinterpret_cast<const uint4*>(my_in + 16u * blk);
      const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
      bool pass = base >= lead && base + 16 <= total && prev >= 0;
      if (pass && relaxed) {
        pass = !pending_sp && prev != '\t';
        uint32_t sp[4], cr[4], tab = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) { sp[q] = canon_eq_mask(w4[q], ' '); cr[q] = canon_eq_mask(w4[q], '\r'); tab |= canon_eq_mask(w4[q], '\t'); }
        uint32_t bad = tab;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const uint32_t nx = ((sp[q] | cr[q]) >> 8) | (q < 3 ? (sp[q + 1] | cr[q + 1]) << 24 : 0u);
          bad |= sp[q] & nx;
        }
        if (prev == ' ') bad |= (sp[0] | cr[0]) & 0x80u;
        pass = pass && bad == 0;
      }
      if (pass) {
        const uint32_t p = (uint32_t)prev;
        const uint32_t o0 = p | (v.x << 8), o1 = __funnelshift_l(v.x, v.y, 8), o2 = __funnelshift_l(v.y, v.z, 8), o3 = __funnelshift_l(v.z, v.w, 8);
        w.emit8((uint64_t)o0 | ((uint64_t)o1 << 32));
        w.emit8((uint64_t)o2 | ((uint64_t)o3 << 32));
        prev = (int)(v.w >> 24);
        continue;
      }
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const uint32_t pos = base + (uint32_t)k;
        const bool valid = pos >= lead && pos < total;
        const uint32_t cur = (w4[k >> 2] >> ((k & 3) * 8)) & 0xffu;
        const bool have = valid && prev >= 0;
        const uint32_t c = (uint32_t)prev;
        if (relaxed) {
          const bool wsp = c == ' ' || c == '\t';
          const bool crlf = c == '\r' && cur == '\n';
          const bool sp_out = have && !wsp && pending_sp && !crlf;
          w.emit(' ', sp_out);
          w.emit(c, have && !wsp);
          pending_sp = have ? wsp : pending_sp;
        } else {
          w.emit(c, have);
        }
        prev = valid ? (int)cur : prev;
      }
    }
    // ---- cooperative write-back of the full 16-byte chunks of every output row
    w.park();
    __syncwarp();
    const uint32_t fill = w.o - w.flushed, nfull = fill >> 4;
#pragma unroll
    for (int g = 0; g < 8; g++) {
      const uint32_t body = 4u * g + (lane >> 3), chunk = lane & 7u;
      const uint32_t bn = __shfl_sync(FULL, nfull, body), bfl = __shfl_sync(FULL, w.flushed, body);
      const uint32_t blo = __shfl_sync(FULL, go_lo, body), bhi = __shfl_sync(FULL, go_hi, body);
      // a row holds at most 9 full chunks (144 bytes); the 9th, when present, goes in a second pass below
      if (chunk < bn) {
        const uint4 v = *reinterpret_cast<const uint4*>(out_rows + body * CANON_OUT_PITCH + 16u * chunk);
        *reinterpret_cast<uint4*>((((uint64_t)bhi << 32) | blo) + bfl + 16u * chunk) = v;
      }
    }
    if (nfull > 8) *reinterpret_cast<uint4*>(gout + w.flushed + 128u) = *reinterpret_cast<const uint4*>(w.row + 128u);
    __syncwarp();
    if (nfull) {   // the remainder moves to the front of the row
      const uint4 r = *reinterpret_cast<const uint4*>(w.row + 16u * nfull);
      *reinterpret_cast<uint4*>(w.row) = r;
      w.flushed += 16u * nfull;
    }
    __syncwarp();
  }
  if (!live) return;
  // ---- tail: the byte without a successor, then what is left in the row, then the post-pass on the global copy
  if (prev >= 0) {
    const uint32_t c = (uint32_t)prev;
    if (relaxed) {
      const bool wsp = c == ' ' || c == '\t';
      w.emit(' ', !wsp && pending_sp);
      w.emit(c, !wsp);
      w.emit(' ', wsp);  // a trailing WSP run is kept as one SP (cfdkim quirk: "abc " -> "abc \r\n")
    } else {
      w.emit(c, true);
    }
  }
  w.park();
  for (uint32_t k = w.flushed; k < w.o; k++) gout[k] = w.row[k - w.flushed];
  uint32_t o = w.o;
  uint8_t* out = gout;
  if (relaxed) {
    while (o >= 4 && out[o - 1] == '\n' && out[o - 2] == '\r' && out[o - 3] == '\n' && out[o - 4] == '\r') o -= 2;
    if (o > 0 && !(o >= 2 && out[o - 2] == '\r' && out[o - 1] == '\n')) { out[o] = '\r'; out[o + 1] = '\n'; o += 2; }
  } else {
    if (n == 0) { out[0] = '\r'; out[1] = '\n'; o = 2; }
    else while (o >= 4 && out[o - 1] == '\n' && out[o - 2] == '\r' && out[o - 3] == '\n' && out[o - 4] == '\r') o -= 2;
  }
  if ((it.flags & 2u) && it.l < o) o = it.l;  // l=: Vec::truncate of the canonical body
  msg_len[it.msg] = o;
}

}  // namespace zkb
