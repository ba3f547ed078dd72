// dfa.cuh — batched dense-DFA regex scan for sm_100a (kernel K3).
//
// Replaces regex_automata::dfa::regex::Regex::find_iter as called by process_regex_parts
// (core/src/regex.rs:32-39): forward unanchored leftmost-first DFA to the match end, anchored
// reverse DFA back to the match start, non-overlapping iteration over the whole haystack (the
// predicate is "exactly one match", so the whole haystack is always consumed).  Search semantics
// follow SURVEY.md A.5: premultiplied state ids, byte classes, match states delayed by one byte,
// an end-of-input class, dead state 0, start state chosen by the byte before the search start.
//
// For body haystacks the kernel fuses remove_quoted_printable_soft_breaks
// (core/src/email.rs:61-86): "=\r\n" triples are skipped on the fly and the haystack is virtually
// zero-padded back to its original length, so no cleaned copy is ever materialised in HBM.  All
// offsets reported are in cleaned coordinates, as in the reference.
//
// Mapping: one thread per (email, pattern); the pattern's forward and reverse tables live in
// shared memory (copied once per CTA, 128-bit copies).  Per byte the work is a dependent chain of
// shared-memory loads (one for DIRECT tables with 256-entry rows, two - class then transition - for
// class-compressed tables); throughput comes from occupancy.  Haystack bytes are fetched 16 at a time
// (one LDG.128 per 16 bytes per lane) into registers; the inner loop is unrolled over the block.
#pragma once
#include "common.cuh"

namespace zkb {

// Device table blob (built by the host from ZDF1, engine.cu: build_dfa_blob):
//   u32[0]=n_states [1]=row stride (n_classes, or 256 for DIRECT tables) [2]=min_match*stride
//   [3]=max_match*stride [4]=flags [5]=elem bytes [18]=1 for DIRECT tables
//   u32[6..17] = start ids (premultiplied): unanchored[6], anchored[6]
//   byte 128: class_map[256]; byte 384: start_map[256]; byte 640: trans (u16 or u32, premultiplied by
//   the row stride); DIRECT tables append eoi[n_states]
#define ZKB_DFA_HDR 640
#define ZKB_DFA_UTF8 2u
#define ZKB_DFA_HAS_EMPTY 4u

// DIRECT tables have one 256-entry row per state (no byte-class indirection: one shared-memory load
// per byte) followed by an end-of-input column eoi[n_states]; compressed tables index rows by byte
// class (two dependent loads per byte) and keep the EOI transition as the last class.
template <typename TT, bool DIRECT>
struct DfaTab {
  const TT* trans;
  const TT* eoi;
  const uint8_t* cmap;
  const uint8_t* smap;
  const uint32_t* hdr;
  uint32_t ncls, min_m, max_m, flags;
  __device__ __forceinline__ void init(const uint8_t* blob) {
    hdr = reinterpret_cast<const uint32_t*>(blob);
    ncls = hdr[1]; min_m = hdr[2]; max_m = hdr[3]; flags = hdr[4];
    cmap = blob + 128; smap = blob + 384;
    trans = reinterpret_cast<const TT*>(blob + ZKB_DFA_HDR);
    eoi = trans + (size_t)hdr[0] * 256;
  }
  __device__ __forceinline__ uint32_t next(uint32_t sid, uint32_t byte) const {
    if (DIRECT) return trans[sid + byte];
    return trans[sid + cmap[byte]];
  }
  __device__ __forceinline__ uint32_t next_eoi(uint32_t sid) const {
    if (DIRECT) return eoi[sid >> 8];
    return trans[sid + ncls - 1];
  }
  // the host numbers match states last (dfa_host.hpp): match <=> sid >= min_m; without match states min_m is
  // above every id
  __device__ __forceinline__ bool is_match(uint32_t sid) const { return sid >= min_m; }
  __device__ __forceinline__ uint32_t start(bool anchored, uint32_t kind) const { return hdr[6 + (anchored ? 6 : 0) + kind]; }
};

// A position in the cleaned stream: c = cleaned index, o = index into the original bytes of the
// next byte to consume (always normalised past soft breaks; o == n once the original is exhausted,
// after which the stream continues with zero padding up to cleaned index n).
struct Cur {
  uint32_t c, o;
};

template <typename TT, bool DIRECT>
struct Searcher {
  DfaTab<TT, DIRECT> f, r;
  const uint8_t* h;   // 16-byte aligned; readable up to the next 16-byte boundary past n
  uint32_t n;
  bool qp;
  uint32_t clen;      // cleaned length of the real bytes; known once a forward scan reached the end
  uint4 win;          // register window: the 16-byte block win_blk of the haystack
  uint32_t win_blk;

  __device__ __forceinline__ uint32_t byte_at(uint32_t o) {
    const uint32_t blk = o >> 4;
    if (blk != win_blk) { win = __ldg(reinterpret_cast<const uint4*>(h) + blk); win_blk = blk; }
    const uint32_t w = (o & 8) ? ((o & 4) ? win.w : win.z) : ((o & 4) ? win.y : win.x);
    return (w >> ((o & 3) * 8)) & 0xffu;
  }
  __device__ __forceinline__ void skip_soft(uint32_t& o) {
    if (qp)
      while (o + 2 < n && byte_at(o) == '=' && byte_at(o + 1) == '\r' && byte_at(o + 2) == '\n') o += 3;
  }
  __device__ __forceinline__ Cur begin() {
    Cur p = {0u, 0u};
    skip_soft(p.o);
    if (p.o >= n) clen = 0;
    return p;
  }
  // consume the byte at p (requires p.c < n) and advance
  __device__ __forceinline__ uint32_t take(Cur& p) {
    uint32_t b = 0;
    if (p.o < n) {
      b = byte_at(p.o);
      p.o++;
      skip_soft(p.o);
      if (p.o >= n) clen = p.c + 1;
    }
    p.c++;
    return b;
  }
  // the cleaned byte just before p (requires p.c > 0); moves p back
  __device__ __forceinline__ uint32_t back(Cur& p) {
    p.c--;
    if (p.o >= n && p.c >= clen) return 0u;  // inside the zero padding
    uint32_t o = p.o - 1;
    if (qp)
      while (o >= 2 && byte_at(o) == '\n' && byte_at(o - 1) == '\r' && byte_at(o - 2) == '=') o -= 3;
    p.o = o;
    return byte_at(o);
  }

  // find_fwd: unanchored leftmost-first from `from` (prev = cleaned byte before it, -1 at text
  // start).  On a match: end = position of the match end, end_byte = byte there (-1 for EOI).
  //
  // The scan walks aligned 16-byte blocks held in registers (one LDG.128 per 16 bytes per lane) and is
  // branch-free inside a block: bytes before the start offset / past the end, and the bytes of "=\r\n" soft
  // breaks (QP), are predicated off instead of taking a slower path, so every lane of a warp executes the same
  // instruction stream whatever its data looks like (a lane that branches on a per-byte event leaves the
  // convergent group; a lane that falls back to a byte-wise path makes its whole warp wait).  The dead state
  // is absorbing (row 0 is all zeros) and never a match, so finishing the block after dying is harmless.
  // One aligned 16-byte block of a forward scan (requires o < n): advances (sid, c, o, skip) and records the
  // last match position seen in the block.
  template <bool QP>
  __device__ __forceinline__ void fwd_block(uint32_t& sid, uint32_t& c, uint32_t& o, uint32_t& skip, bool& have, Cur& end, int& end_byte) {
    const uint4* hb = reinterpret_cast<const uint4*>(h);
    const uint32_t base = o & ~15u;
    const uint32_t kend = (n - base) < 16u ? (n - base) : 16u;
    const uint32_t mask = ((1u << kend) - 1u) & ~((1u << (o & 15u)) - 1u);   // bytes of this block in range
    const uint4 v = __ldg(hb + (base >> 4));
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
    // bytes to drop: the tail of a soft break that started in the previous block, plus every "=\r\n"
    // starting in this one.  Only blocks that contain '=' pay for the search (a short, rarely taken branch;
    // the 16-step transition loop below is common to all lanes).
    uint32_t drop = skip;
    if (QP) {
      const uint32_t e = 0x3d3d3d3du;
      const uint32_t x0 = v.x ^ e, x1 = v.y ^ e, x2 = v.z ^ e, x3 = v.w ^ e;
      const uint32_t z = ((x0 - 0x01010101u) & ~x0) | ((x1 - 0x01010101u) & ~x1) | ((x2 - 0x01010101u) & ~x2) | ((x3 - 0x01010101u) & ~x3);
      if (z & 0x80808080u) {
        uint32_t la = 0;   // the two bytes after the block (look-ahead for k = 14, 15)
        if (base + 16 < n) la = __ldg(reinterpret_cast<const uint32_t*>(h + base + 16));
        const uint32_t w5[5] = {v.x, v.y, v.z, v.w, la};
#pragma unroll
        for (int k = 0; k < 16; k++) {
          const uint32_t b0 = (w5[k >> 2] >> ((k & 3) * 8)) & 0xffu;
          const uint32_t b1 = (w5[(k + 1) >> 2] >> (((k + 1) & 3) * 8)) & 0xffu;
          const uint32_t b2 = (w5[(k + 2) >> 2] >> (((k + 2) & 3) * 8)) & 0xffu;
          const bool sb = b0 == '=' && b1 == '\r' && b2 == '\n' && base + k + 2 < n && ((mask >> k) & 1u);
          drop |= sb ? (7u << k) : 0u;
        }
      }
    }
    skip = drop >> 16;                          // a soft break at k = 14 / 15 spills into the next block
    const uint32_t cons = mask & ~drop;         // bytes the DFA consumes
    int mk = -1;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const uint32_t b = __byte_perm(w4[k >> 2], 0u, 0x4440u + (uint32_t)(k & 3));   // one PRMT per byte
      const bool act = (cons >> k) & 1u;
      const uint32_t ns = f.next(sid, b);
      sid = act ? ns : sid;
      mk = (act && f.is_match(sid)) ? k : mk;
    }
    if (mk >= 0) {
      have = true;
      end.c = c + (uint32_t)__popc(cons & ((1u << mk) - 1u));
      end.o = base + (uint32_t)mk;
      const uint32_t w = (mk & 8) ? ((mk & 4) ? v.w : v.z) : ((mk & 4) ? v.y : v.x);
      end_byte = (int)((w >> ((mk & 3) * 8)) & 0xffu);
    }
    c += (uint32_t)__popc(cons);
    o = base + 16u;
  }
  // What follows the real bytes of a forward scan: virtual zero padding up to the original length (only bodies
  // that had soft breaks get here with c < n), then the end-of-input transition.
  __device__ __forceinline__ void fwd_tail(uint32_t sid, uint32_t c, uint32_t o, bool& have, Cur& end, int& end_byte) {
    while (c < n && sid != 0 && o >= n) {
      sid = f.next(sid, 0u);
      if (f.is_match(sid)) { have = true; end.c = c; end.o = n; end_byte = 0; }
      c++;
    }
    if (sid == 0) return;  // dead: the search ends with the last recorded match
    sid = f.next_eoi(sid);
    if (f.is_match(sid)) { have = true; end.c = n; end.o = n; end_byte = -1; }
  }
  template <bool QP>
  __device__ bool fwd_impl(Cur from, int prev, Cur& end, int& end_byte) {
    if (from.c > n) return false;
    uint32_t sid = f.start(false, prev < 0 ? 2u : f.smap[prev]);
    if (sid == 0) return false;
    bool have = false;
    uint32_t c = from.c, o = from.o, skip = 0;
    while (o < n && sid != 0) fwd_block<QP>(sid, c, o, skip, have, end, end_byte);
    if (from.o < n && o >= n) clen = c;   // crossed the end of the real bytes: the cleaned length is known
    fwd_tail(sid, c, o, have, end, end_byte);
    return have;
  }
  __device__ __forceinline__ bool fwd(Cur from, int prev, Cur& end, int& end_byte) {
    return qp ? fwd_impl<true>(from, prev, end, end_byte) : fwd_impl<false>(from, prev, end, end_byte);
  }
  // find_rev: anchored, over cleaned [from.c, end.c); ms = leftmost start of a match ending at end
  __device__ bool rev(Cur from, int prev_of_from, Cur end, int end_byte, uint32_t& ms) {
    uint32_t sid = r.start(true, end_byte < 0 ? 2u : r.smap[end_byte]);
    if (sid == 0) return false;
    bool have = false;
    Cur p = end;
    while (p.c > from.c && sid != 0) {
      uint32_t b = back(p);
      sid = r.next(sid, b);
      if (r.is_match(sid)) { have = true; ms = p.c + 1; }
    }
    if (sid == 0) return have;
    sid = prev_of_from < 0 ? r.next_eoi(sid) : r.next(sid, (uint32_t)prev_of_from);
    if (r.is_match(sid)) { have = true; ms = from.c; }
    return have;
  }
  // dfa::regex::Regex::find over the span [from, n)
  __device__ bool find(Cur from, int prev, uint32_t& ms, Cur& me, int& me_byte, bool& panic) {
    Cur e = from;
    int eb = -1;
    if (!fwd(from, prev, e, eb)) return false;
    if ((f.flags & ZKB_DFA_UTF8) && (f.flags & ZKB_DFA_HAS_EMPTY)) {
      // util::empty::skip_splits_fwd: an end inside a UTF-8 sequence is not allowed; retry from
      // successive start offsets (on a copy: the caller's span start stays where it was)
      Cur st = from;
      int pv = prev;
      while (eb >= 0 && (int8_t)eb < -0x40) {
        if (st.c >= n) return false;
        pv = (int)take(st);
        if (!fwd(st, pv, e, eb)) return false;
      }
    }
    me = e; me_byte = eb;
    if (e.c == from.c) { ms = e.c; return true; }
    if (!rev(from, prev, e, eb, ms)) { panic = true; return false; }
    return true;
  }
  // find_iter: count all non-overlapping matches, keep the first span
  __device__ void run(uint32_t& count, uint32_t& fs, uint32_t& fe, bool& panic) {
    count = 0; fs = 0; fe = 0; panic = false;
    clen = 0xffffffffu;
    win_blk = 0xffffffffu;
    Cur pos = begin();
    int prev = -1;
    uint32_t last_end = 0xffffffffu;
    for (;;) {
      uint32_t ms; Cur me; int mb;
      if (!find(pos, prev, ms, me, mb, panic)) break;
      if (ms == me.c && me.c == last_end) {  // handle_overlapping_empty_match
        if (pos.c >= n) break;
        prev = (int)take(pos);
        if (!find(pos, prev, ms, me, mb, panic)) break;
      }
      if (count == 0) { fs = ms; fe = me.c; }
      count++;
      if (me.c > pos.c) {  // next search starts at the match end; prev = cleaned byte before it
        Cur t = me;
        prev = (int)back(t);
        pos = me;
      }
      last_end = me.c;
    }
  }
  // find_iter for patterns that cannot match the empty string, as ONE loop over blocks: a lane whose search ends
  // (dead state or end of input) books the match and starts its next search inside the same loop instead of
  // leaving it, so the lanes of a warp - whose matches sit at different offsets - keep executing the block code
  // together.  (With one loop per search the warp runs max(first search) + max(second search) block steps,
  // about twice what each lane needs.)  The reverse scans that only produce match starts are deferred: the first
  // and second match are resolved after the loop, later ones (already "not exactly one") on the spot, which keeps
  // the count identical to running find() per match even for inconsistent forward / reverse tables.
  // Returns false when an empty match shows up after all (tables whose flags lie): the caller then uses run().
  template <bool QP>
  __device__ bool run_flat(uint32_t& count, uint32_t& fs, uint32_t& fe, bool& panic) {
    count = 0; fs = 0; fe = 0; panic = false;
    clen = 0xffffffffu;
    win_blk = 0xffffffffu;
    Cur pos = begin();
    int prev = -1;
    const Cur from0 = pos;
    Cur e1 = pos, e2 = pos, from2 = pos;
    int b1 = -1, b2 = -1, prev2 = -1;
    uint32_t sid = f.start(false, 2u), c = pos.c, o = pos.o, skip = 0, from_o = pos.o;
    bool have = false, done = sid == 0;
    Cur end = pos;
    int end_byte = -1;
    while (!done) {
      if (o < n && sid != 0) { fwd_block<QP>(sid, c, o, skip, have, end, end_byte); continue; }
      // this search is over
      if (from_o < n && o >= n) clen = c;
      fwd_tail(sid, c, o, have, end, end_byte);
      if (!have) break;
      if (end.c == pos.c) return false;   // empty match: not this path
      if (count == 0) { e1 = end; b1 = end_byte; }
      else if (count == 1) { e2 = end; b2 = end_byte; from2 = pos; prev2 = prev; }
      else { uint32_t ms; if (!rev(pos, prev, end, end_byte, ms)) { panic = true; break; } }
      count++;
      // next search: from the match end, look-behind = the cleaned byte before it
      Cur t = end;
      prev = (int)back(t);
      pos = end;
      sid = f.start(false, f.smap[prev]);
      c = pos.c; o = pos.o; skip = 0; from_o = pos.o; have = false;
      done = sid == 0;
    }
    if (count >= 1) {
      uint32_t ms;
      if (!rev(from0, -1, e1, b1, ms)) { count = 0; panic = true; return true; }
      fs = ms; fe = e1.c;
    }
    if (count >= 2) {
      uint32_t ms;
      if (!rev(from2, prev2, e2, b2, ms)) { count = 1; panic = true; }
    }
    return true;
  }
};

#ifdef ZKB_HOST_EMU
#define ZKB_DYN_SMEM(name) static uint8_t name[232448] __attribute__((aligned(16)))
#else
#define ZKB_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#endif

// Copies both tables into shared memory (SMEM instantiations; 128-bit copies, blobs are 16-byte padded
// by the host).  SMEM is a template parameter so that the table pointers are derived unconditionally
// from the shared array and the scan uses LDS rather than generic loads.
template <bool SMEM>
__device__ __forceinline__ void dfa_stage_tables(uint8_t* smem, const uint8_t* __restrict__ fwd_blob, uint32_t fwd_bytes,
                                                 const uint8_t* __restrict__ rev_blob, uint32_t rev_bytes,
                                                 const uint8_t*& fb, const uint8_t*& rb) {
  if (SMEM) {
    const uint32_t fpad = (fwd_bytes + 15u) & ~15u, rpad = (rev_bytes + 15u) & ~15u;
    for (uint32_t i = threadIdx.x * 16; i < fpad; i += blockDim.x * 16)
      *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(fwd_blob + i);
    for (uint32_t i = threadIdx.x * 16; i < rpad; i += blockDim.x * 16)
      *reinterpret_cast<uint4*>(smem + fpad + i) = *reinterpret_cast<const uint4*>(rev_blob + i);
    __syncthreads();
    fb = smem; rb = smem + fpad;
  } else {
    fb = fwd_blob; rb = rev_blob;
  }
}

template <typename TT, bool DIRECT>
__device__ __forceinline__ uint4 dfa_scan_one(const uint8_t* fb, const uint8_t* rb, const uint8_t* hay, uint32_t n, int qp) {
  Searcher<TT, DIRECT> s;
  s.f.init(fb); s.r.init(rb);
  s.h = hay; s.n = n; s.qp = qp != 0;
  uint32_t count, fs, fe; bool panic;
  bool flat = false;
  if (!(s.f.flags & ZKB_DFA_HAS_EMPTY)) flat = s.qp ? s.template run_flat<true>(count, fs, fe, panic) : s.template run_flat<false>(count, fs, fe, panic);
  if (!flat) s.run(count, fs, fe, panic);
  return make_uint4(count, fs, fe, panic ? 1u : 0u);
}

// out[slot] = (match_count, first start, first end, reverse-search-failed flag)
template <typename TT, bool DIRECT, bool SMEM>
__global__ void __launch_bounds__(128)
dfa_scan_kernel(const uint8_t* __restrict__ arena, const DfaItem* __restrict__ items,
                uint32_t n_items, const uint32_t* __restrict__ msg_len, const uint8_t* __restrict__ fwd_blob,
                uint32_t fwd_bytes, const uint8_t* __restrict__ rev_blob, uint32_t rev_bytes, int qp,
                uint4* __restrict__ out) {
  ZKB_DYN_SMEM(smem);
  const uint8_t *fb, *rb;
  dfa_stage_tables<SMEM>(smem, fwd_blob, fwd_bytes, rev_blob, rev_bytes, fb, rb);
  uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_items) return;
  DfaItem it = items[idx];
  out[it.out_slot] = dfa_scan_one<TT, DIRECT>(fb, rb, arena + it.hay_off, msg_len[it.msg], qp);
}

// Engine form: items are interleaved per email (2*j = header preimage, 2*j+1 = canonical body);
// `which` selects the haystack, the result of regex part `pi` of P goes to out[email * P + pi].
template <typename TT, bool DIRECT, bool SMEM>
__global__ void __launch_bounds__(128)
dfa_scan_strided(const uint8_t* __restrict__ arena, const DfaItem* __restrict__ items, uint32_t n_emails,
                 const uint32_t* __restrict__ msg_len, uint32_t which, uint32_t P, uint32_t pi,
                 const uint8_t* __restrict__ fwd_blob, uint32_t fwd_bytes,
                 const uint8_t* __restrict__ rev_blob, uint32_t rev_bytes, int qp,
                 uint4* __restrict__ out) {
  ZKB_DYN_SMEM(smem);
  const uint8_t *fb, *rb;
  dfa_stage_tables<SMEM>(smem, fwd_blob, fwd_bytes, rev_blob, rev_bytes, fb, rb);
  uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_emails) return;
  DfaItem it = items[2 * idx + which];
  out[(size_t)it.out_slot * P + pi] = dfa_scan_one<TT, DIRECT>(fb, rb, arena + it.hay_off, msg_len[it.msg], qp);
}

}  // namespace zkb
