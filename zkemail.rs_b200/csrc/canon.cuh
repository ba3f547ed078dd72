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
// 8 output bytes).  No data-dependent branches in the byte loop: the lanes of a warp stay converged.
#pragma once
#include "common.cuh"

namespace zkb {

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

}  // namespace zkb
