// frontend_warp.cuh — the device-side DKIM front end, warp-cooperative form (kernel K-1 of DESIGN.md; sm_100a).
//
// Same contract as frontend.cuh (which stays as the scalar twin the emulated tests compare against): for one raw message
//   mailparse header split -> first DKIM-Signature header of from_domain -> tag list -> required tags, v=1, d= ==
//   from_domain, c=, a=rsa-sha256, optional i= q= x= l= -> signed-header selection per h= (bottom-up, repeated names
//   walk upward) -> relaxed/simple header canonicalisation -> the b=-blanked DKIM-Signature header without its final
//   CRLF (the header-hash preimage, written into the arena slot the SHA-256 kernel reads) -> base64 of bh= (8 digest
//   words) and b= (little-endian signature limbs for the RSA kernel).
// Reference behaviour: cfdkim::verify_email_with_key / validate_header / select_headers (core/src/email.rs:31-33;
// SURVEY.md Appendix A.2).
//
// Mapping: ONE WARP PER MESSAGE.  The header block is staged once into shared memory with coalesced 16-byte loads
// (every lane a different 16 bytes; the search for the terminating CRLF CRLF rides on the same pass); every later
// pass works on 32 consecutive bytes per step, one per lane, with __ballot_sync masks for the delimiters (LF / ':' /
// ';' / FWS), population counts for stream compaction (relaxed canonicalisation, h= names, base64 characters) and
// lane-per-item work for the short irregular pieces (one header key per lane, one tag per lane, one signature limb
// per lane).  All control flow is warp-uniform, so the warp never diverges on message content — the scalar form ran
// 12.8 of 32 lanes per instruction (ncu, round 1) and one global load per byte window.
//
// Acceptance is NARROWER than the scalar form (which was already conservative): on top of its rules the header block
// must fit FE_HB_CAP bytes, use CRLF line ends throughout (no bare CR / LF), every header line must have a colon, the
// tag list must parse to its end (an optional trailing ';'), <= FE_MAXT tags, <= FE_B64_CAP base64 characters in b=.
// Everything else is FE_FALLBACK: the engine re-runs that message through the host front end (dkim_host.hpp), which
// implements every path.  On what it accepts this code must be — and is tested to be — byte-identical to the host.
#pragma once
#include "common.cuh"
#include "frontend.cuh"   // character classes, FeOut / FeIn users, the scalar twin

namespace zkb {

#define FE_HB_CAP 4096     // header block bytes staged per message
#define FE_MAXT 32         // tags in a DKIM-Signature header
#define FE_B64_CAP 704     // base64 characters of b= (4096-bit signature: 684)
#define FE_WARPS 4         // warps (messages) per CTA

struct __align__(16) FeWarpSmem {
  uint8_t hdr_raw[FE_HB_CAP + 80];   // the staged bytes in the 16-byte phase of global memory: message byte i = hdr_raw[lead + i]
  uint16_t hstart[FE_MAXH + 2];   // header starts, then the end of the header block (offset of the blank line)
  uint16_t hcolon[FE_MAXH];
  uint16_t hvs[FE_MAXH];          // value start (spaces after the colon skipped)
  uint32_t hhash[FE_MAXH];        // hash of the lower-cased key
  uint16_t hkl[FE_MAXH];          // key length for relaxed canonicalisation; bit 15: the key has a non-ASCII byte
  uint16_t seg[FE_MAXT + 2];      // tag-list segment boundaries (offsets of ';' inside the value, then its length)
  uint16_t tvoff[FE_MAXT], tvlen[FE_MAXT];
  uint8_t tcode[FE_MAXT];
  uint8_t hbuf[256];              // h= value with FWS removed
  uint16_t name_s[FE_MAXN + 1], name_e[FE_MAXN + 1];
  uint32_t name_hash[FE_MAXN];
  int16_t hit_of[FE_MAXN];
  uint8_t b64[FE_B64_CAP + 8];    // base64 characters of the tag being decoded, FWS removed
  uint8_t dec[FE_B64_CAP / 4 * 3 + 8];   // their decoded bytes (written by the validation pass, one quad per lane)
};

// Character classes and base64 values, one table lookup each (shared memory on the device: lanes index it with
// different bytes, which constant memory would serialise).
struct FeLut { uint8_t cls[256]; uint8_t b64[256]; };
enum { C_FWS = 1, C_VAL = 2, C_ALPHA = 4, C_ALNUM = 8, C_UPPER = 16, C_WSP = 32 };
__device__ __forceinline__ void fe_lut_fill(FeLut* t, uint32_t tid, uint32_t nthreads) {
  for (uint32_t c = tid; c < 256; c += nthreads) {
    uint32_t k = 0;
    if (fe_fws(c)) k |= C_FWS;
    if (fe_valchar(c)) k |= C_VAL;
    if (fe_alpha(c)) k |= C_ALPHA;
    if (fe_alnum_(c)) k |= C_ALNUM;
    if (c - 'A' < 26u) k |= C_UPPER;
    if (c == ' ' || c == '\t') k |= C_WSP;
    t->cls[c] = (uint8_t)k;
    const int v = fe_b64(c);
    t->b64[c] = v < 0 ? 0xFFu : (uint8_t)v;
  }
}

// FNV-1a of a literal and its little-endian packing into words, evaluated by the compiler
__host__ __device__ constexpr uint32_t fe_fnv(const char* s, uint32_t h = 2166136261u) { return *s ? fe_fnv(s + 1, (h ^ (uint32_t)(uint8_t)*s) * 16777619u) : h; }
__host__ __device__ constexpr uint32_t fe_len(const char* s) { return *s ? 1u + fe_len(s + 1) : 0u; }
__host__ __device__ constexpr uint32_t fe_word(const char* s, uint32_t k) {   // bytes 4k .. 4k+3 of the literal, zero padded
  uint32_t w = 0, n = fe_len(s);
  for (uint32_t b = 0; b < 4; b++) if (4 * k + b < n) w |= (uint32_t)(uint8_t)s[4 * k + b] << (8 * b);
  return w;
}

// tag codes (bit = 1 << code in `seen`)
enum { T_V = 0, T_A, T_B, T_BH, T_D, T_H, T_C, T_S, T_I, T_Q, T_X, T_L, T_OTHER = 31 };

// The whole front end for one message, executed by one warp.  Outputs as fe_process (frontend.cuh).
__device__ inline void fe_process_warp(FeWarpSmem* sm, const FeLut* lut, const uint8_t* raw, uint32_t n, const uint8_t* dom, uint32_t dom_len, uint32_t k,
                                       uint32_t limbs, uint8_t* pre, uint32_t* sigw, FeOut& out, uint32_t& body_l, bool allow_skip,
                                       long long now) {
  const unsigned FULL = 0xffffffffu;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t ltm = (1u << lane) - 1u;
  auto CLS = [&](uint32_t c) -> uint32_t { return lut->cls[c & 0xffu]; };
  auto LOW = [&](uint32_t c) -> uint32_t { return c + ((uint32_t)(lut->cls[c & 0xffu] & C_UPPER) << 1); };
  auto FWSQ = [&](uint32_t c) -> bool { return (lut->cls[c & 0xffu] & C_FWS) != 0; };
  auto B64 = [&](uint32_t c) -> int { const uint32_t v = lut->b64[c & 0xffu]; return v == 0xFFu ? -1 : (int)v; };
  body_l = 0;
  out.flags = 0; out.body_off = 0; out.body_len = 0; out.pre_len = 0;
  for (int i = 0; i < 8; i++) out.bh[i] = 0;
  for (uint32_t i = lane; i < limbs; i += 32) sigw[i] = 0;
#define FE_FAIL(code) do { out.flags = (code); return; } while (0)

  // ---------------------------------------------------------------- 1. stage + find the end of the header block
  if (n < 4) FE_FAIL(FE_FALLBACK);
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(raw) & ~(uintptr_t)15;
  const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(raw) - a0);
  const uint32_t want = n < FE_HB_CAP ? n : FE_HB_CAP;
  uint32_t body_off = 0, checked = 0;
  const uint8_t* hdr = sm->hdr_raw + lead;              // message byte i
  for (uint32_t base = 0; base < want + lead && !body_off; base += 512) {
    // every lane one aligned 16-byte block, stored as it is (the shared-memory copy keeps the phase of global memory)
    const uint32_t pos = base + lane * 16;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (pos < want + lead) v = __ldg(reinterpret_cast<const uint4*>(a0 + pos));
    *reinterpret_cast<uint4*>(sm->hdr_raw + pos) = v;
    __syncwarp();
    const uint32_t hi = (base + 512 < want + lead ? base + 512 : want + lead);   // raw bytes staged so far (incl. the lead)
    // CRLF CRLF among the candidate positions (raw coordinates) whose four bytes are staged: 16 candidates per lane out
    // of five aligned shared-memory words; candidates that still miss bytes are looked at again in the next step
    {
      const uint32_t c0 = checked + lane * 16;
      uint32_t first = 0xffffffffu;
      if (c0 + 4 <= hi) {
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(sm->hdr_raw + c0);
        uint32_t x[5];
#pragma unroll
        for (int q = 0; q < 5; q++) x[q] = wp[q];
#pragma unroll
        for (int q = 15; q >= 0; q--) {
          const uint32_t four = (q & 3) ? __funnelshift_r(x[q >> 2], x[(q >> 2) + 1], 8 * (q & 3)) : x[q >> 2];
          if (four == 0x0a0d0a0du && c0 + q >= lead && c0 + q + 4 <= hi) first = c0 + q;
        }
      }
      const unsigned hit = __ballot_sync(FULL, first != 0xffffffffu);
      if (hit) body_off = __shfl_sync(FULL, first, __ffs((int)hit) - 1) + 4 - lead;
      checked = hi >= 3 ? ((hi - 3) & ~15u) : 0u;     // every candidate below this had its four bytes
    }
  }
  if (!body_off) FE_FAIL(FE_FALLBACK);                  // no CRLF CRLF within FE_HB_CAP bytes
  const uint32_t hend = body_off - 2;                   // end of the last header line (its CRLF included)
  const uint32_t c0 = hdr[0];
  if (c0 == '\r' || c0 == '\n' || c0 == ' ' || c0 == '\t') FE_FAIL(FE_FALLBACK);   // empty / malformed first line: host decides

  // ---------------------------------------------------------------- 2. header starts; CRLF discipline
  uint32_t nh = 0;
  {
    // three delimiter masks per 32 bytes; everything else is the same bit arithmetic on every lane.  Every LF must
    // follow a CR and every CR must be followed by a LF (mailparse splits on LF alone; bare ones go to the host).
    unsigned bad = 0, prev_lf = 1u /* a header starts at offset 0 */, prev_cr = 0u;
    for (uint32_t base = 0; base < hend; base += 32) {
      const uint32_t i = base + lane;
      const uint32_t c = i < hend ? hdr[i] : 'x';
      const unsigned lf = __ballot_sync(FULL, c == '\n'), cr = __ballot_sync(FULL, c == '\r');
      const unsigned ws = __ballot_sync(FULL, c == ' ' || c == '\t');
      const unsigned valid = hend - base >= 32 ? FULL : ((1u << (hend - base)) - 1u);
      bad |= lf & ~((cr << 1) | prev_cr);                 // LF without CR
      bad |= (cr & 0x7fffffffu) & ~(lf >> 1);              // CR without LF (bit 31 is judged with the next step)
      bad |= prev_cr & ~lf & 1u;
      const unsigned starts = ((lf << 1) | prev_lf) & ~ws & valid;
      if ((starts >> lane) & 1u) { const uint32_t r = nh + __popc(starts & ltm); if (r < FE_MAXH) sm->hstart[r] = (uint16_t)i; }
      nh += __popc(starts);
      prev_lf = lf >> 31; prev_cr = cr >> 31;
    }
    if (bad) FE_FAIL(FE_FALLBACK);
    if (nh > FE_MAXH) FE_FAIL(FE_FALLBACK);
    if (lane == 0) sm->hstart[nh] = (uint16_t)hend;
    __syncwarp();
  }
  out.body_off = body_off; out.body_len = n - body_off;

  // ---------------------------------------------------------------- 3. one header per lane: colon, value start, key hash
  unsigned sig_mask[2] = {0u, 0u};
  {
    bool bad = false;
    constexpr uint32_t kSigHash = fe_fnv("dkim-signature");
    for (uint32_t hb = 0; hb < nh; hb += 32) {
      const uint32_t h = hb + lane;
      bool is_sig = false;
      if (h < nh) {
        const uint32_t s = sm->hstart[h], e = sm->hstart[h + 1];
        uint32_t p = s, hv = 2166136261u, hi8 = 0;
        for (; p < e; p++) {
          const uint32_t c = hdr[p];
          if (c == ':' || c == '\n') break;
          hv = (hv ^ LOW(c)) * 16777619u;
          hi8 |= c;
        }
        if (p >= e || hdr[p] != ':') bad = true;
        else {
          sm->hcolon[h] = (uint16_t)p;
          sm->hhash[h] = hv;
          uint32_t q = p + 1;
          while (q < e && hdr[q] == ' ') q++;
          sm->hvs[h] = (uint16_t)q;
          uint32_t kl = p - s;                             // relaxed key: trailing SP / 0x09..0x0d dropped
          while (kl > 0 && (hdr[s + kl - 1] == ' ' || (uint32_t)(hdr[s + kl - 1] - 9u) <= 4u)) kl--;
          sm->hkl[h] = (uint16_t)(kl | ((hi8 & 0x80u) ? 0x8000u : 0u));
          is_sig = p - s == 14 && hv == kSigHash;
        }
      }
      sig_mask[hb >> 5] = __ballot_sync(FULL, is_sig);
    }
    if (__ballot_sync(FULL, bad)) FE_FAIL(FE_FALLBACK);
    __syncwarp();
    // the hashes matched: confirm the 14 bytes (a colliding key is just another header)
    for (int w = 0; w < 2; w++) {
      unsigned m = sig_mask[w];
      while (m) {
        const uint32_t h = 32u * (uint32_t)w + (uint32_t)__ffs((int)m) - 1u;
        m &= m - 1;
        const char* lit = "dkim-signature";
        const bool eq = lane >= 14 || LOW(hdr[sm->hstart[h] + lane]) == (uint32_t)(uint8_t)lit[lane];
        if (__ballot_sync(FULL, !eq)) sig_mask[w] &= ~(1u << (h & 31u));
      }
    }
  }
  const uint32_t n_sigs = (uint32_t)(__popc(sig_mask[0]) + __popc(sig_mask[1]));
  if (n_sigs == 0 || n_sigs > 8) FE_FAIL(FE_FALLBACK);

  // value of header h: [hvs, hstart[h+1] - 2)
  auto val_off = [&](uint32_t h) -> uint32_t { return sm->hvs[h]; };
  auto val_len = [&](uint32_t h) -> uint32_t { return (uint32_t)sm->hstart[h + 1] - 2u - sm->hvs[h]; };

  // ---------------------------------------------------------------- 4. DKIM-Signature headers, top to bottom
  uint32_t so = 0, sn = 0, seen = 0;
  FeVal tv{0, 0}, ta{0, 0}, tb{0, 0}, tbh{0, 0}, td{0, 0}, th{0, 0}, tc{0, 0}, ti{0, 0}, tq{0, 0}, tx{0, 0}, tl{0, 0};
  // A short tag value with its FWS removed, packed into four little-endian words (every lane gets the same words);
  // len = number of characters (99 when the value cannot be one of the literals it is compared with)
  struct Packed { uint32_t w[4]; uint32_t len; };
  auto pack_value = [&](FeVal v) -> Packed {
    Packed pk;
    pk.w[0] = pk.w[1] = pk.w[2] = pk.w[3] = 0; pk.len = 99;
    if (v.len > 32) return pk;
    uint32_t c = 0;
    bool keep = false;
    if (lane < v.len) { c = hdr[so + v.off + lane]; keep = !FWSQ(c); }
    const unsigned m = __ballot_sync(FULL, keep);
    const uint32_t r = __popc(m & ltm);
#pragma unroll
    for (uint32_t q = 0; q < 4; q++) pk.w[q] = __reduce_or_sync(FULL, (keep && (r >> 2) == q) ? c << (8 * (r & 3)) : 0u);
    pk.len = __popc(m) <= 16 ? (uint32_t)__popc(m) : 99u;
    return pk;
  };
#define FE_IS(pk, lit) ((pk).len == fe_len(lit) && (pk).w[0] == fe_word(lit, 0) && (pk).w[1] == fe_word(lit, 1) && \
                        (pk).w[2] == fe_word(lit, 2) && (pk).w[3] == fe_word(lit, 3))
  // 0: a well-formed rsa-sha256 signature of from_domain; 2: well-formed, another domain; 1: anything else.  Warp-uniform.
  auto parse_sig = [&](uint32_t idx) -> int {
    so = val_off(idx); sn = val_len(idx);
    // ---- byte classes of the whole value; ';' positions are the segment boundaries
    uint32_t nseg = 0;
    {
      bool bad = false;
      for (uint32_t base = 0; base < sn; base += 32) {
        const uint32_t i = base + lane;
        bool semi = false;
        if (i < sn) {
          const uint32_t c = hdr[so + i];
          semi = c == ';';
          bad = bad || !(semi || (CLS(c) & (C_FWS | C_VAL)));       // control characters, DEL, non-ASCII
        }
        const unsigned m = __ballot_sync(FULL, semi);
        if (semi) { const uint32_t r = nseg + __popc(m & ltm); if (r < FE_MAXT) sm->seg[r] = (uint16_t)i; }
        nseg += __popc(m);
      }
      if (__ballot_sync(FULL, bad)) return 1;
    }
    if (nseg >= FE_MAXT) return 1;
    if (lane == 0) sm->seg[nseg] = (uint16_t)sn;
    __syncwarp();
    const uint32_t ntag = nseg + 1;                                  // segments; the last one may be empty (trailing ';')
    // ---- one segment per lane
    bool bad = false, empty = false;
    uint32_t code = T_OTHER, voff = 0, vlen = 0;
    if (lane < ntag) {
      uint32_t a = lane == 0 ? 0u : (uint32_t)sm->seg[lane - 1] + 1u, b = sm->seg[lane];
      uint32_t p = a;
      while (p < b && FWSQ(hdr[so + p])) p++;
      if (p >= b) empty = true;
      else if (!(CLS(hdr[so + p]) & C_ALPHA)) bad = true;
      else {
        const uint32_t name_off = p;
        while (p < b && (CLS(hdr[so + p]) & C_ALNUM)) p++;
        const uint32_t name_len = p - name_off;
        while (p < b && FWSQ(hdr[so + p])) p++;
        if (p >= b || hdr[so + p] != '=') bad = true;
        else {
          p++;
          while (p < b && FWSQ(hdr[so + p])) p++;
          voff = p;
          uint32_t e = b;
          while (e > p && FWSQ(hdr[so + e - 1])) e--;
          vlen = e - p;                                              // valchar runs joined by FWS (classes checked above)
          const uint32_t n0 = hdr[so + name_off], n1 = name_len > 1 ? hdr[so + name_off + 1] : 0u;
          if (name_len == 1) {
            switch (n0) {
              case 'v': code = T_V; break; case 'a': code = T_A; break; case 'b': code = T_B; break; case 'd': code = T_D; break;
              case 'h': code = T_H; break; case 'c': code = T_C; break; case 's': code = T_S; break; case 'i': code = T_I; break;
              case 'q': code = T_Q; break; case 'x': code = T_X; break; case 'l': code = T_L; break; default: break;
            }
          } else if (name_len == 2 && n0 == 'b' && n1 == 'h') code = T_BH;
        }
      }
    }
    // an empty segment is fine only as the last one; the first segment must be a tag
    if (__ballot_sync(FULL, bad || (empty && (lane + 1 != ntag || lane == 0)))) return 1;
    seen = 0;
    {
      uint32_t dup = 0;
      for (uint32_t c = 0; c <= T_L; c++) {
        const unsigned m = __ballot_sync(FULL, lane < ntag && !empty && code == c);
        if (__popc(m) > 1) dup = 1;
        if (m) {
          seen |= 1u << c;
          const int src = __ffs((int)m) - 1;
          FeVal v;
          v.off = __shfl_sync(FULL, voff, src); v.len = __shfl_sync(FULL, vlen, src);
          switch (c) {
            case T_V: tv = v; break; case T_A: ta = v; break; case T_B: tb = v; break; case T_BH: tbh = v; break;
            case T_D: td = v; break; case T_H: th = v; break; case T_C: tc = v; break; case T_I: ti = v; break;
            case T_Q: tq = v; break; case T_X: tx = v; break; case T_L: tl = v; break; default: break;
          }
        }
      }
      if (dup) return 1;
    }
    const uint32_t need = (1u << T_V) | (1u << T_A) | (1u << T_B) | (1u << T_BH) | (1u << T_D) | (1u << T_H) | (1u << T_S);
    if ((seen & need) != need) return 1;
    // ---- the short literal values: packed once, compared as words
    {
      const Packed pv = pack_value(tv), pa = pack_value(ta);
      if (!FE_IS(pv, "1") || !FE_IS(pa, "rsa-sha256") || tb.len == 0) return 1;
      if (seen & (1u << T_Q)) { const Packed pq = pack_value(tq); if (!FE_IS(pq, "dns/txt")) return 1; }
    }
    int rc = 0;
    uint32_t lval = 0;
    if (lane == 0) {
      if (seen & (1u << T_I)) {                                      // i= must end with the d= value (bytes, FWS removed)
        uint32_t li = 0, ld = 0;
        for (uint32_t i = 0; i < ti.len; i++) if (!FWSQ(hdr[so + ti.off + i])) li++;
        for (uint32_t i = 0; i < td.len; i++) if (!FWSQ(hdr[so + td.off + i])) ld++;
        if (li < ld) rc = 1;
        uint32_t a = ti.len, b = td.len;
        for (uint32_t m = 0; m < ld && rc == 0; m++) {
          do { a--; } while (FWSQ(hdr[so + ti.off + a]));
          do { b--; } while (FWSQ(hdr[so + td.off + b]));
          if (hdr[so + ti.off + a] != hdr[so + td.off + b]) rc = 1;
        }
      }
      if (rc == 0 && (seen & (1u << T_X))) {                         // x=: plain decimal clearly in the future of `now`
        long long x = 0;
        uint32_t digits = 0;
        for (uint32_t i = 0; i < tx.len && rc == 0; i++) {
          const uint32_t c = hdr[so + tx.off + i];
          if (FWSQ(c)) continue;
          if (c < '0' || c > '9' || ++digits > 17) rc = 1;
          else x = x * 10 + (long long)(c - '0');
        }
        if (rc == 0 && (digits == 0 || now + 2 > x + 15 * 60)) rc = 1;
      }
      if (rc == 0 && (seen & (1u << T_L))) {                         // l=: [+]digits
        uint32_t digits = 0, plus = 0;
        unsigned long long v = 0;
        for (uint32_t i = 0; i < tl.len && rc == 0; i++) {
          const uint32_t c = hdr[so + tl.off + i];
          if (FWSQ(c)) continue;
          if (c == '+' && digits == 0 && plus == 0) { plus = 1; continue; }
          if (c < '0' || c > '9' || ++digits > 18) rc = 1;
          else v = v * 10 + (unsigned long long)(c - '0');
        }
        if (rc == 0 && digits == 0) rc = 1;
        lval = v > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)v;
      }
    }
    rc = __shfl_sync(FULL, rc, 0);
    if (rc == 0) {                                                   // d= == from_domain (ASCII case-insensitive, FWS removed)
      uint32_t j = 0;
      bool differ = false;
      for (uint32_t base = 0; base < td.len; base += 32) {
        const uint32_t i = base + lane;
        uint32_t c = 0;
        bool keep = false;
        if (i < td.len) { c = hdr[so + td.off + i]; keep = !FWSQ(c); }
        const unsigned m = __ballot_sync(FULL, keep);
        const uint32_t r = j + __popc(m & ltm);
        if (keep) differ = differ || r >= dom_len || LOW(c) != LOW(dom[r]);
        j += __popc(m);
      }
      if (__ballot_sync(FULL, differ) || j != dom_len) rc = 2;       // a signature of another domain: the reference skips it
    }
    body_l = __shfl_sync(FULL, lval, 0);
    return rc;
  };

  int sig_idx = -1;
  for (uint32_t q = 0; q < n_sigs && sig_idx < 0; q++) {
    // q-th set bit of the signature masks
    uint32_t idx, left = q;
    if (left < (uint32_t)__popc(sig_mask[0])) { unsigned m = sig_mask[0]; while (left--) m &= m - 1; idx = (uint32_t)__ffs((int)m) - 1; }
    else { left -= __popc(sig_mask[0]); unsigned m = sig_mask[1]; while (left--) m &= m - 1; idx = 32u + (uint32_t)__ffs((int)m) - 1; }
    const int r = parse_sig(idx);
    if (r == 1 || (r == 2 && !allow_skip)) FE_FAIL(FE_FALLBACK);
    if (r == 0) sig_idx = (int)idx;
  }
  if (sig_idx < 0) FE_FAIL(FE_FALLBACK);
  const uint32_t multi = (n_sigs > 1 ? FE_MULTI : 0u) | ((seen & (1u << T_L)) ? FE_HAS_L : 0u);
  bool hr = false, br = false;
  if (seen & (1u << T_C)) {
    const Packed pc = pack_value(tc);
    if (FE_IS(pc, "relaxed/relaxed")) { hr = true; br = true; }
    else if (FE_IS(pc, "simple/simple") || FE_IS(pc, "simple")) { hr = false; br = false; }
    else if (FE_IS(pc, "relaxed/simple") || FE_IS(pc, "relaxed")) { hr = true; br = false; }
    else if (FE_IS(pc, "simple/relaxed")) { hr = false; br = true; }
    else FE_FAIL(FE_FALLBACK);
  }

  // ---------------------------------------------------------------- 5. h=: names, bottom-up selection
  uint32_t hl = 0;
  for (uint32_t base = 0; base < th.len; base += 32) {
    const uint32_t i = base + lane;
    uint32_t c = 0;
    bool keep = false;
    if (i < th.len) { c = hdr[so + th.off + i]; keep = !FWSQ(c); }
    const unsigned m = __ballot_sync(FULL, keep);
    const uint32_t r = hl + __popc(m & ltm);
    if (keep && r < sizeof sm->hbuf) sm->hbuf[r] = (uint8_t)c;
    hl += __popc(m);
  }
  if (hl > sizeof sm->hbuf) FE_FAIL(FE_FALLBACK);
  __syncwarp();
  uint32_t nn = 0;
  {
    // name boundaries: a name ends at every ':' and at the end; empty names are dropped
    uint32_t prev_colon = 0xffffffffu;   // index of the last ':' seen so far (carried across steps), -1 initially
    for (uint32_t base = 0; base <= hl; base += 32) {
      const uint32_t i = base + lane;
      const bool delim = i <= hl && (i == hl || sm->hbuf[i] == ':');
      const unsigned m = __ballot_sync(FULL, delim);
      // start of the name that ends at i = one past the previous delimiter
      uint32_t a;
      const unsigned lower = m & ltm;
      if (lower) a = base + (31u - (uint32_t)__clz((int)lower)) + 1u;
      else a = prev_colon + 1u;          // 0xffffffff + 1 = 0
      const bool name = delim && i > a;
      const unsigned nm = __ballot_sync(FULL, name);
      if (name) { const uint32_t r = nn + __popc(nm & ltm); if (r < FE_MAXN) { sm->name_s[r] = (uint16_t)a; sm->name_e[r] = (uint16_t)i; } }
      nn += __popc(nm);
      if (m) prev_colon = base + (31u - (uint32_t)__clz((int)m));
    }
  }
  if (nn > FE_MAXN) FE_FAIL(FE_FALLBACK);
  __syncwarp();
  bool has_from = false;
  if (lane < nn) {
    const uint32_t a = sm->name_s[lane], e = sm->name_e[lane];
    uint32_t hv = 2166136261u;
    for (uint32_t t = a; t < e; t++) hv = (hv ^ LOW(sm->hbuf[t])) * 16777619u;
    sm->name_hash[lane] = hv;
    has_from = e - a == 4 && LOW(sm->hbuf[a]) == 'f' && LOW(sm->hbuf[a + 1]) == 'r' && LOW(sm->hbuf[a + 2]) == 'o' && LOW(sm->hbuf[a + 3]) == 'm';
  }
  if (!__ballot_sync(FULL, has_from)) FE_FAIL(FE_FALLBACK);
  __syncwarp();
  for (uint32_t j = 0; j < nn; j++) {
    const uint32_t nl = (uint32_t)sm->name_e[j] - sm->name_s[j], nhash = sm->name_hash[j];
    // cursor: the latest earlier occurrence of the same name
    bool same = false;
    if (lane < j) same = (uint32_t)sm->name_e[lane] - sm->name_s[lane] == nl && sm->name_hash[lane] == nhash;
    const unsigned sm_mask = __ballot_sync(FULL, same);
    int start = (int)nh;
    if (sm_mask) {
      const int i = 31 - __clz((int)sm_mask);
      // the hashes matched: confirm the bytes (a collision is not worth a wrong cursor)
      bool eq = true;
      if (lane < nl) eq = LOW(sm->hbuf[sm->name_s[i] + lane]) == LOW(sm->hbuf[sm->name_s[j] + lane]);
      for (uint32_t t = 32 + lane; t < nl; t += 32) eq = eq && LOW(sm->hbuf[sm->name_s[i] + t]) == LOW(sm->hbuf[sm->name_s[j] + t]);
      if (__ballot_sync(FULL, !eq)) FE_FAIL(FE_FALLBACK);
      const int hprev = sm->hit_of[i];
      start = hprev >= 0 ? hprev : 0;
    }
    int hit = -1;
    for (int hb = (int)((nh - 1) & ~31u); hb >= 0 && hit < 0; hb -= 32) {
      const int x = hb + (int)lane;
      bool m = false;
      if (x < start && x < (int)nh) m = (uint32_t)sm->hcolon[x] - sm->hstart[x] == nl && sm->hhash[x] == nhash;
      const unsigned mm = __ballot_sync(FULL, m);
      if (mm) hit = hb + 31 - __clz((int)mm);
    }
    if (hit >= 0) {
      bool eq = true;
      for (uint32_t t = lane; t < nl; t += 32) eq = eq && LOW(hdr[sm->hstart[hit] + t]) == LOW(sm->hbuf[sm->name_s[j] + t]);
      if (__ballot_sync(FULL, !eq)) FE_FAIL(FE_FALLBACK);
    }
    if (lane == 0) sm->hit_of[j] = (int16_t)hit;
    __syncwarp();
  }

  // ---------------------------------------------------------------- 6. the preimage
  uint32_t o = 0;
  bool overflow = false;
  // cooperative copies into pre[]
  auto put_lit = [&](const char* s, uint32_t len) {
    if (lane < len) { if (o + lane < FE_PRE_CAP) pre[o + lane] = (uint8_t)s[lane]; else overflow = true; }
    o += len;
  };
  auto put_raw = [&](uint32_t off, uint32_t len, bool lower) {
    for (uint32_t t = lane; t < len; t += 32) {
      const uint32_t c = hdr[off + t];
      if (o + t < FE_PRE_CAP) pre[o + t] = (uint8_t)(lower ? LOW(c) : c); else overflow = true;
    }
    o += len;
  };
  // relaxed value: CR / LF dropped (they only occur as CRLF here), WSP runs collapsed to one SP, leading and trailing
  // WSP dropped; bytes in [skip_lo, skip_hi) are absent.  A WSP byte is never written by itself: the non-WSP byte that
  // follows a run writes the SP in front of itself, so a trailing run leaves nothing behind (no byte is written twice —
  // lanes of a warp have no write order among each other).  State carried between calls: prev_sp, rv_start.
  bool prev_sp = true;
  uint32_t rv_start = 0;
  auto put_relaxed = [&](uint32_t off, uint32_t len, uint32_t skip_lo, uint32_t skip_hi) {
    for (uint32_t base = 0; base < len; base += 32) {
      const uint32_t i = base + lane;
      uint32_t c = 0;
      bool present = false, wsp = false;
      if (i < len && !(i >= skip_lo && i < skip_hi)) {
        c = hdr[off + i];
        present = c != '\r' && c != '\n';
        wsp = c == ' ' || c == '\t';
      }
      const unsigned pm = __ballot_sync(FULL, present), wm = __ballot_sync(FULL, present && wsp);
      const unsigned nw = pm & ~wm;                           // bytes that are written
      const unsigned lower = pm & ltm;
      const bool prev_w = lower ? ((wm >> (31 - __clz((int)lower))) & 1u) != 0 : prev_sp;
      // a SP goes in front of a non-WSP byte that follows a WSP run, unless nothing of this value has been written yet
      const bool started = o > rv_start || (nw & ltm) != 0;
      const unsigned spm = __ballot_sync(FULL, present && !wsp && prev_w && started);
      if ((nw >> lane) & 1u) {
        uint32_t at = o + __popc(nw & ltm) + __popc(spm & ltm);
        if ((spm >> lane) & 1u) { if (at < FE_PRE_CAP) pre[at] = ' '; else overflow = true; at++; }
        if (at < FE_PRE_CAP) pre[at] = (uint8_t)c; else overflow = true;
      }
      o += __popc(nw) + __popc(spm);
      if (pm) prev_sp = ((wm >> (31 - __clz((int)pm))) & 1u) != 0;
    }
  };
  auto finish_relaxed = [&]() { put_lit("\r\n", 2); };
  for (uint32_t j = 0; j < nn; j++) {
    const int hit = sm->hit_of[j];
    if (hit < 0) continue;
    const uint32_t ks = sm->hstart[hit], kl0 = (uint32_t)sm->hcolon[hit] - ks;
    if (sm->hkl[hit] & 0x8000u) FE_FAIL(FE_FALLBACK);            // non-ASCII byte in a selected key
    if (hr) {
      const uint32_t kl = sm->hkl[hit] & 0x7fffu;
      put_raw(ks, kl, true);
      put_lit(":", 1);
      prev_sp = true; rv_start = o;
      put_relaxed(val_off((uint32_t)hit), val_len((uint32_t)hit), 0xffffffffu, 0xffffffffu);
      finish_relaxed();
    } else {
      put_raw(ks, kl0, false);
      put_lit(": ", 2);
      put_raw(val_off((uint32_t)hit), val_len((uint32_t)hit), false);
      put_lit("\r\n", 2);
    }
  }
  // ---- the signature header with the raw b= text removed (value.replace(raw_b, "")); another occurrence of that
  // text anywhere in the value is left to the host
  {
    bool dup = false;
    const uint32_t b0 = hdr[so + tb.off];
    for (uint32_t p = lane; p + tb.len <= sn; p += 32) {
      if (p == tb.off || hdr[so + p] != b0) continue;
      bool same = true;
      for (uint32_t t = 1; same && t < tb.len; t++) same = hdr[so + p + t] == hdr[so + tb.off + t];
      dup = dup || same;
    }
    if (__ballot_sync(FULL, dup)) FE_FAIL(FE_FALLBACK);
  }
  if (hr) {
    put_lit("dkim-signature:", 15);
    prev_sp = true; rv_start = o;
    // the b= value (most of the header) is absent: only the bytes before and after it are walked, with the state carried
    put_relaxed(so, tb.off, 0xffffffffu, 0xffffffffu);
    put_relaxed(so + tb.off + tb.len, sn - tb.off - tb.len, 0xffffffffu, 0xffffffffu);
    finish_relaxed();
  } else {
    put_lit("DKIM-Signature: ", 16);
    put_raw(so, tb.off, false);
    put_raw(so + tb.off + tb.len, sn - tb.off - tb.len, false);
    put_lit("\r\n", 2);
  }
  if (__ballot_sync(FULL, overflow) || o > FE_PRE_CAP || o < 2) FE_FAIL(FE_FALLBACK);
  out.pre_len = o - 2;   // final CRLF dropped
  uint32_t flags = (hr ? FE_HDR_RELAXED : 0u) | (br ? FE_BODY_RELAXED : 0u) | multi;

  // ---------------------------------------------------------------- 7. base64 of bh= and b=
  // compacts the characters of a tag value (FWS removed) into sm->b64; returns their number (or ~0u when too many)
  auto compact = [&](FeVal v) -> uint32_t {
    uint32_t cnt = 0;
    for (uint32_t base = 0; base < v.len; base += 32) {
      const uint32_t i = base + lane;
      uint32_t c = 0;
      bool keep = false;
      if (i < v.len) { c = hdr[so + v.off + i]; keep = !FWSQ(c); }
      const unsigned m = __ballot_sync(FULL, keep);
      const uint32_t r = cnt + __popc(m & ltm);
      if (keep && r < FE_B64_CAP) sm->b64[r] = (uint8_t)c;
      cnt += __popc(m);
    }
    __syncwarp();
    return cnt <= FE_B64_CAP ? cnt : 0xffffffffu;
  };
  // strict STANDARD base64 over sm->b64[0..nchars): returns the decoded length or -1 (syntax).  One quad per lane.
  auto validate = [&](uint32_t nchars) -> int {
    if (nchars % 4) return -1;
    const uint32_t nq = nchars / 4;
    bool bad = false;
    uint32_t pad = 0;
    for (uint32_t qb = 0; qb < nq; qb += 32) {
      const uint32_t q = qb + lane;
      if (q < nq) {
        const uint32_t c0 = sm->b64[4 * q], c1 = sm->b64[4 * q + 1], c2 = sm->b64[4 * q + 2], c3 = sm->b64[4 * q + 3];
        const int a = B64(c0), b = B64(c1), c = B64(c2), d = B64(c3);
        // the decoded bytes go to sm->dec as the quad is judged (bytes of a padded or broken quad are never read)
        const uint32_t ua = (uint32_t)a & 63u, ub = (uint32_t)b & 63u, uc = (uint32_t)c & 63u, ud = (uint32_t)d & 63u;
        sm->dec[3 * q] = (uint8_t)((ua << 2) | (ub >> 4));
        sm->dec[3 * q + 1] = (uint8_t)((ub << 4) | (uc >> 2));
        sm->dec[3 * q + 2] = (uint8_t)((uc << 6) | ud);
        if (a < 0 || b < 0) bad = true;
        else if (!(c >= 0 && d >= 0)) {
          if (q + 1 != nq) bad = true;
          else if (c2 == '=' && c3 == '=') { if (b & 15) bad = true; else pad = 2; }
          else if (c3 == '=' && c >= 0) { if (c & 3) bad = true; else pad = 1; }
          else bad = true;
        }
      }
    }
    __syncwarp();
    if (__ballot_sync(FULL, bad)) return -1;
    const unsigned pm = __ballot_sync(FULL, pad != 0);
    const uint32_t p = pm ? __shfl_sync(FULL, pad, __ffs((int)pm) - 1) : 0u;
    return (int)(3 * nq - p);
  };
  auto byte_at = [&](uint32_t i) -> uint32_t { return sm->dec[i]; };   // decoded byte i (from the start)
  {
    const uint32_t cnt = compact(tbh);
    int dl = -1;
    if (cnt == 44) dl = validate(44);
    if (dl == 32) {
      flags |= FE_BH_VALID;
      uint32_t wv = 0;
      if (lane < 8) wv = (byte_at(4 * lane) << 24) | (byte_at(4 * lane + 1) << 16) | (byte_at(4 * lane + 2) << 8) | byte_at(4 * lane + 3);
      for (int i = 0; i < 8; i++) out.bh[i] = __shfl_sync(FULL, wv, i);
    }
    __syncwarp();
  }
  {
    const uint32_t cnt = compact(tb);
    if (cnt == 0xffffffffu) FE_FAIL(FE_FALLBACK);
    const int sl = validate(cnt);
    if (sl < 0) flags |= FE_SIG_SYNTAX;
    else if ((uint32_t)sl != k || k > 4 * limbs) flags |= FE_SIG_BADLEN;
    else {
      // limb L holds the big-endian bytes sl-1-4L .. sl-4-4L (missing high bytes are zero)
      for (uint32_t L = lane; L < limbs; L += 32) {
        uint32_t wv = 0;
#pragma unroll
        for (uint32_t bi = 0; bi < 4; bi++) {
          const uint32_t pos = 4 * L + bi;              // byte index from the least significant end
          if (pos < (uint32_t)sl) wv |= byte_at((uint32_t)sl - 1 - pos) << (8 * bi);
        }
        sigw[L] = wv;
      }
    }
  }
  out.flags = flags;
#undef FE_FAIL
#undef FE_IS
}

#ifdef ZKB_HOST_EMU
#define ZKB_FE_SMEM(name) static FeWarpSmem name[FE_WARPS]; static FeLut name##_lut
#else
#define ZKB_FE_SMEM(name) __shared__ FeWarpSmem name[FE_WARPS]; __shared__ FeLut name##_lut
#endif

// One warp per message.  Also emits the body's CanonItem (an empty one for fallback / error messages).
__global__ void __launch_bounds__(FE_WARPS * 32)
frontend_warp_kernel(const uint8_t* __restrict__ span, const FeIn* __restrict__ in, uint32_t n, uint8_t* __restrict__ arena,
                     const uint64_t* __restrict__ msg_off, uint32_t* __restrict__ msg_len, uint32_t* __restrict__ sig_arena,
                     uint32_t* __restrict__ cand_bh, CanonItem* __restrict__ canon, FeOut* __restrict__ out, int allow_skip,
                     long long now) {
  ZKB_FE_SMEM(smem);
  fe_lut_fill(&smem_lut, threadIdx.x, FE_WARPS * 32);
  __syncthreads();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t idx = blockIdx.x * FE_WARPS + warp;
  if (idx >= n) return;
  const FeIn fi = in[idx];
  FeOut fo;
  uint32_t body_l = 0;
  fe_process_warp(&smem[warp], &smem_lut, span + fi.raw_off, fi.raw_len, arena + fi.dom_off, fi.dom_len, fi.k, fi.limbs, arena + msg_off[fi.pre_msg],
                  sig_arena + fi.sig_word_off, fo, body_l, allow_skip != 0, now);
  if (lane != 0) return;
  const bool live = (fo.flags & (FE_FALLBACK | FE_MAIL_PARSE)) == 0;
  msg_len[fi.pre_msg] = live ? fo.pre_len : 0u;
  CanonItem ci;
  ci.raw_off = fi.raw_off + fo.body_off;
  ci.raw_len = live ? fo.body_len : 0u;
  ci.msg = fi.body_msg;
  ci.flags = ((fo.flags & FE_BODY_RELAXED) ? 1u : 0u) | ((fo.flags & FE_HAS_L) ? 2u : 0u);
  ci.l = body_l; ci.pad[0] = ci.pad[1] = 0;
  canon[idx] = ci;
  for (int i = 0; i < 8; i++) cand_bh[(size_t)fi.cand * 8 + i] = fo.bh[i];
  out[idx] = fo;
}

}  // namespace zkb
