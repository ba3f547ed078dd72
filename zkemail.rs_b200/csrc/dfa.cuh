// dfa.cuh — batched dense-DFA regex scan for sm_100a (kernel K3).
//
// Replaces regex_automata::dfa::regex::Regex::find_iter as called by process_regex_parts
// (core/src/regex.rs:32-39): forward unanchored leftmost-first DFA to the match end, anchored
// reverse DFA back to the match start, non-overlapping iteration over the whole haystack
// (the predicate is "exactly one match", so the whole haystack is always consumed).
// Search semantics follow SURVEY.md A.5: premultiplied state ids, byte classes, match states
// delayed by one byte, an end-of-input class, dead state 0, start state chosen by the byte
// before the search start.
//
// For body haystacks the kernel fuses remove_quoted_printable_soft_breaks
// (core/src/email.rs:61-86): "=\r\n" triples are skipped on the fly and the haystack is
// virtually zero-padded back to its original length, so no cleaned copy is ever materialised.
// All offsets reported are in cleaned coordinates, as in the reference.
//
// Mapping: one thread per (email, pattern); the pattern's forward and reverse tables live in
// shared memory (copied once per CTA).  The per-byte work is a dependent chain of two shared
// loads (class, transition); throughput comes from occupancy.
#pragma once
#include "common.cuh"

namespace zkb {

// Device table blob (built by the host from ZDF1, see engine.cu:build_dfa_blob):
//   u32[0]=n_states [1]=n_classes [2]=min_match*ncls [3]=max_match*ncls [4]=flags [5]=elem bytes
//   u32[6..17] = start ids (premultiplied): unanchored[6], anchored[6]
//   byte 128: class_map[256]; byte 384: start_map[256]; byte 640: trans (u16 or u32)
#define ZKB_DFA_HDR 640
#define ZKB_DFA_UTF8 2u
#define ZKB_DFA_HAS_EMPTY 4u

template <typename TT>
struct DfaView {
  const TT* trans;
  const uint8_t* cmap;
  const uint8_t* smap;
  const uint32_t* hdr;
  uint32_t ncls, min_m, max_m, flags;
  __device__ __forceinline__ void init(const uint8_t* blob) {
    hdr = reinterpret_cast<const uint32_t*>(blob);
    ncls = hdr[1]; min_m = hdr[2]; max_m = hdr[3]; flags = hdr[4];
    cmap = blob + 128; smap = blob + 384;
    trans = reinterpret_cast<const TT*>(blob + ZKB_DFA_HDR);
  }
  __device__ __forceinline__ uint32_t next(uint32_t sid, uint32_t byte) const {
    return trans[sid + cmap[byte]];
  }
  __device__ __forceinline__ uint32_t next_eoi(uint32_t sid) const { return trans[sid + ncls - 1]; }
  __device__ __forceinline__ bool is_match(uint32_t sid) const { return sid - min_m <= max_m - min_m && min_m <= max_m; }
  __device__ __forceinline__ uint32_t start(bool anchored, uint32_t kind) const {
    return hdr[6 + (anchored ? 6 : 0) + kind];
  }
};

// Cleaned view of a haystack: with qp, "=\r\n" is skipped and the tail is zero padded to n.
struct Hay {
  const uint8_t* h;
  uint32_t n;
  bool qp;
  __device__ __forceinline__ bool soft_break_at(uint32_t o) const {
    return qp && o + 2 < n && h[o] == '=' && h[o + 1] == '\r' && h[o + 2] == '\n';
  }
  __device__ __forceinline__ bool soft_break_ending_at(uint32_t o) const {  // h[o-2..o]
    return qp && o >= 2 && h[o] == '\n' && h[o - 1] == '\r' && h[o - 2] == '=';
  }
};

struct Pos {       // a position in the cleaned stream
  uint32_t c;      // cleaned index
  uint32_t o;      // original index of the byte at cleaned index c (== n inside the zero padding)
};

template <typename TT>
struct Searcher {
  DfaView<TT> f, r;
  Hay hay;

  // cleaned byte at p (p.c < n); advances p to the next cleaned position
  __device__ __forceinline__ uint32_t take(Pos& p) const {
    while (p.o < hay.n && hay.soft_break_at(p.o)) p.o += 3;
    uint32_t b = (p.o < hay.n) ? hay.h[p.o] : 0u;
    if (p.o < hay.n) p.o++;
    p.c++;
    return b;
  }
  // forward search from `from` (prev = cleaned byte before it, or -1 at start of text).
  // Returns true with the match end (cleaned index, its Pos and the byte at the end or -1 for EOI).
  __device__ bool fwd(Pos from, int prev, Pos& end, int& end_byte) const {
    if (from.c > hay.n) return false;
    uint32_t sid = f.start(false, prev < 0 ? 2u : f.smap[prev]);
    if (sid == 0) return false;
    bool have = false;
    Pos p = from;
    while (p.c < hay.n) {
      Pos at = p;
      // normalise `at.o` so that it points at the byte actually consumed
      while (at.o < hay.n && hay.soft_break_at(at.o)) at.o += 3;
      p = at;
      uint32_t b = take(p);
      sid = f.next(sid, b);
      if (sid <= f.max_m) {
        if (sid == 0) return have;
        if (f.is_match(sid)) { have = true; end = at; end_byte = (int)b; }
      }
    }
    sid = f.next_eoi(sid);
    if (f.is_match(sid)) { have = true; end = p; end_byte = -1; }
    return have;
  }
  // anchored reverse search over cleaned [from.c, end.c); returns the leftmost start (cleaned idx)
  __device__ bool rev(Pos from, int prev_of_from, Pos end, int end_byte, uint32_t& ms) const {
    uint32_t sid = r.start(true, end_byte < 0 ? 2u : r.smap[end_byte]);
    if (sid == 0) return false;
    bool have = false;
    uint32_t c = end.c;
    // number of virtual zero bytes between the true cleaned length and end.c: they sit at
    // cleaned indices where the original is exhausted (end.o == n)
    uint32_t o = end.o;  // original index one past the last byte to consume (exclusive)
    // cleaned length of the real bytes = c - zeros; we discover zeros by walking: while the
    // original cursor is at n and we still have cleaned positions that are padding.
    // A position is padding iff its index >= cleaned_len, where cleaned_len = c_at(o==n first).
    // We recover cleaned_len lazily: count real bytes in [from.o, n) only when end.o == n.
    uint32_t real_len_from = 0;  // cleaned count of real bytes in [from.o, end.o)
    if (end.o >= hay.n && hay.qp) {
      uint32_t oo = from.o;
      while (oo < hay.n) {
        if (hay.soft_break_at(oo)) { oo += 3; continue; }
        real_len_from++; oo++;
      }
    } else real_len_from = 0xffffffffu;
    while (c > from.c) {
      uint32_t b;
      if (hay.qp && end.o >= hay.n && (c - from.c) > real_len_from) { b = 0; }  // padding byte
      else {
        o--;
        while (hay.soft_break_ending_at(o)) o -= 3;
        b = hay.h[o];
      }
      c--;
      sid = r.next(sid, b);
      if (sid <= r.max_m) {
        if (sid == 0) return have ? (true) : false;
        if (r.is_match(sid)) { have = true; ms = c + 1; }
      }
    }
    sid = (from.c == 0 || prev_of_from < 0) ? r.next_eoi(sid) : r.next(sid, (uint32_t)prev_of_from);
    if (r.is_match(sid)) { have = true; ms = from.c; }
    return have;
  }
  __device__ __forceinline__ bool char_boundary(const Pos& e, int end_byte) const {
    if (e.c >= hay.n || end_byte < 0) return true;
    return (int8_t)end_byte >= -0x40;
  }
  // dfa::regex::Regex::find in the span [from, n)
  __device__ bool find(Pos from, int prev, uint32_t& ms, Pos& me, int& me_byte, bool& panic) const {
    Pos e; int eb = -1;
    if (!fwd(from, prev, e, eb)) return false;
    if ((f.flags & ZKB_DFA_UTF8) && (f.flags & ZKB_DFA_HAS_EMPTY)) {  // skip_splits_fwd
      // works on a clone of the input: the caller's span start is unchanged afterwards
      Pos st = from; int pv = prev;
      while (!char_boundary(e, eb)) {
        if (st.c >= hay.n) return false;
        pv = (int)take(st);
        if (!fwd(st, pv, e, eb)) return false;
      }
    }
    me = e; me_byte = eb;
    if (e.c == from.c) { ms = e.c; return true; }
    uint32_t s;
    if (!rev(from, prev, e, eb, s)) { panic = true; return false; }
    ms = s;
    return true;
  }
  // find_iter: count all non-overlapping matches, keep the first span
  __device__ void run(uint32_t& count, uint32_t& fs, uint32_t& fe, bool& panic) const {
    count = 0; fs = 0; fe = 0; panic = false;
    Pos pos = {0u, 0u};
    int prev = -1;
    uint32_t last_end = 0xffffffffu;
    for (;;) {
      uint32_t ms; Pos me; int mb;
      if (!find(pos, prev, ms, me, mb, panic)) break;
      if (ms == me.c && me.c == last_end) {  // handle_overlapping_empty_match
        if (pos.c >= hay.n) break;             // start would move past len+1
        prev = (int)take(pos);
        if (!find(pos, prev, ms, me, mb, panic)) break;
      }
      if (count == 0) { fs = ms; fe = me.c; }
      count++;
      // next search starts at the match end; prev = cleaned byte before it
      if (me.c > pos.c) {
        // walk from pos to me to learn the previous byte (cheap: only the matched region)
        Pos w = pos; int pv = prev;
        while (w.c < me.c) pv = (int)take(w);
        prev = pv; pos = w;
      }
      last_end = me.c;
      if (count > hay.n + 1) break;  // safety net
    }
  }
};

template <typename TT>
__global__ void __launch_bounds__(128)
dfa_scan_kernel(const uint8_t* __restrict__ arena, const DfaItem* __restrict__ items,
                uint32_t n_items, const uint8_t* __restrict__ fwd_blob, uint32_t fwd_bytes,
                const uint8_t* __restrict__ rev_blob, uint32_t rev_bytes, int use_smem, int qp,
                uint4* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t smem[];
  const uint8_t* fb = fwd_blob;
  const uint8_t* rb = rev_blob;
  if (use_smem) {
    uint32_t fpad = (fwd_bytes + 15u) & ~15u;
    for (uint32_t i = threadIdx.x * 16; i < fpad; i += blockDim.x * 16)
      *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(fwd_blob + i);
    uint32_t rpad = (rev_bytes + 15u) & ~15u;
    for (uint32_t i = threadIdx.x * 16; i < rpad; i += blockDim.x * 16)
      *reinterpret_cast<uint4*>(smem + fpad + i) = *reinterpret_cast<const uint4*>(rev_blob + i);
    __syncthreads();
    fb = smem; rb = smem + fpad;
  }
  uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_items) return;
  DfaItem it = items[idx];
  Searcher<TT> s;
  s.f.init(fb); s.r.init(rb);
  s.hay.h = arena + it.hay_off; s.hay.n = it.hay_len; s.hay.qp = qp != 0;
  uint32_t count, fs, fe; bool panic;
  s.run(count, fs, fe, panic);
  out[it.out_slot] = make_uint4(count, fs, fe, panic ? 1u : 0u);
}

}  // namespace zkb
