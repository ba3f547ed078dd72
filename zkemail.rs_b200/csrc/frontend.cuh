// frontend.cuh — device-side DKIM front end for sm_100a (kernel K-1; the step after SURVEY.md §8f
// rank 1): for messages that sit in registered host memory the raw bytes are DMA'd as they are and one
// thread per message does, on the device, what dkim_host.hpp does on the host threads for the common
// well-formed case:
//   mailparse header split -> the first DKIM-Signature header of from_domain -> tag list -> required tags, v=1,
//   d= == from_domain, c=, a=rsa-sha256 -> signed-header selection per h= (bottom-up, repeated names walk
//   upward) -> relaxed/simple header canonicalisation -> the b=-blanked DKIM-Signature header without its
//   final CRLF (the header-hash preimage, written into the arena slot the SHA-256 kernel reads) ->
//   base64 of bh= (8 digest words) and b= (little-endian signature limbs for the RSA kernel).
// Reference behaviour: cfdkim::verify_email_with_key / validate_header / select_headers
// (core/src/email.rs:31-33).
//
// Several DKIM-Signature headers are accepted when every header before the first one of from_domain is a
// well-formed signature of another domain (the reference skips those) and, with regex parts, only when the first
// header is that candidate (the haystacks come from the first valid header); such a message is final on the device
// only if the candidate verifies (FE_MULTI), otherwise the host front end takes it.
// Anything that is not the plain passing shape — no DKIM-Signature header, non-ASCII
// bytes in the signature header or a selected key, duplicate tags, i= / q= / x= / l= tags that do not pass, any validation
// error, unknown c=/a=, domain mismatch, a header block not ending in CRLF CRLF, more than FE_MAXH
// headers or FE_MAXN names in h=, an oversized preimage — sets FE_FALLBACK and the engine re-runs that
// message through the host front end, which implements every error path.  The device code therefore
// has to be exact only on the inputs it accepts, and is conservative in what it accepts.
// (Only a malformed header block is decided here: FE_MAIL_PARSE, the reference's parse_mail panic.)
#pragma once
#include "common.cuh"

namespace zkb {

struct FeHdr { uint32_t key_off, key_len, val_off, val_len; };

// Byte reader over one message with a 16-byte register window: sequential scans cost one LDG.128 per 16 bytes
// instead of one L1 sector request per byte and lane (ncu: 155 M sectors for 36 MB of headers without it).
struct FeRd {
  uintptr_t a0;
  uint32_t lead, blk;
  uint4 win;
  __device__ __forceinline__ void init(const uint8_t* p) {
    a0 = reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)15;
    lead = (uint32_t)(reinterpret_cast<uintptr_t>(p) - a0);
    blk = 0xffffffffu;
    win = make_uint4(0u, 0u, 0u, 0u);
  }
  __device__ __forceinline__ uint32_t operator()(uint32_t i) {
    const uint32_t pos = i + lead, b = pos >> 4;
    if (b != blk) { win = __ldg(reinterpret_cast<const uint4*>(a0) + b); blk = b; }
    const uint32_t w = (pos & 8) ? ((pos & 4) ? win.w : win.z) : ((pos & 4) ? win.y : win.x);
    return (w >> ((pos & 3) * 8)) & 0xffu;
  }
};

__device__ __forceinline__ bool fe_fws(uint32_t c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n'; }
__device__ __forceinline__ bool fe_valchar(uint32_t c) { return (c >= 0x21 && c <= 0x3A) || (c >= 0x3C && c <= 0x7E); }
__device__ __forceinline__ bool fe_alpha(uint32_t c) { return ((c | 32u) - 'a') < 26u; }
__device__ __forceinline__ bool fe_alnum_(uint32_t c) { return fe_alpha(c) || (c - '0') < 10u || c == '_'; }
__device__ __forceinline__ uint32_t fe_lower(uint32_t c) { return (c - 'A') < 26u ? c + 32u : c; }

// mailparse::parse_headers (same rules as dkim_host.hpp: parse_headers).  0 ok, 1 parse error, 2 too many
__device__ inline int fe_parse_headers(FeRd& d, uint32_t n, FeHdr* hs, uint32_t& nh, uint32_t& body_off) {
  nh = 0;
  uint32_t ix = 0;
  while (ix < n) {
    const uint32_t c0 = d(ix);
    if (c0 == '\n') { ix++; break; }
    if (c0 == '\r') {
      if (ix + 1 < n && d(ix + 1) == '\n') { ix += 2; break; }
      return 1;
    }
    if (c0 == ' ') return 1;
    uint32_t p = ix;
    while (p < n && d(p) != ':' && d(p) != '\n') p++;
    if (nh >= FE_MAXH) return 2;
    FeHdr h;
    h.key_off = ix;
    if (p >= n) { h.key_len = 0; h.val_off = ix; h.val_len = 0; hs[nh++] = h; ix = n; break; }
    if (d(p) == '\n') { h.key_len = p - ix; h.val_off = p; h.val_len = 0; hs[nh++] = h; ix = p + 1; continue; }
    h.key_len = p - ix;
    p++;
    while (p < n && d(p) == ' ') p++;
    const uint32_t vs = p;
    uint32_t ve = p;
    for (;;) {  // value: until a LF not followed by SP/TAB; end = one past the last byte that is not CR/LF
      uint32_t q = p;
      while (q < n && d(q) != '\n') q++;
      uint32_t e = q;
      while (e > p && d(e - 1) == '\r') e--;
      if (e > p) ve = e;
      if (q >= n) { p = n; break; }
      p = q + 1;
      if (p < n && (d(p) == ' ' || d(p) == '\t')) continue;
      break;
    }
    h.val_off = vs; h.val_len = ve - vs;
    hs[nh++] = h;
    ix = p;
  }
  body_off = ix;
  return 0;
}

// a tag value as a slice of the header value, FWS inside kept
struct FeVal { uint32_t off, len; };

// compares the value with FWS removed against a literal
__device__ inline bool fe_val_is(FeRd& R, uint32_t so, FeVal v, const char* lit) {
  uint32_t j = 0;
  for (uint32_t i = 0; i < v.len; i++) {
    const uint32_t c = R(so + v.off + i);
    if (fe_fws(c)) continue;
    if (lit[j] == 0 || (uint32_t)(uint8_t)lit[j] != c) return false;
    j++;
  }
  return lit[j] == 0;
}

// relaxed header value streamer (the device twin of dkim_host.hpp: RelaxedValue)
struct FeRelaxed {
  uint8_t* out;
  uint32_t o, start, cap;
  bool prev_sp, pending_cr, overflow;
  __device__ __forceinline__ void init(uint8_t* dst, uint32_t at, uint32_t capacity) {
    out = dst; o = at; start = at; cap = capacity; prev_sp = true; pending_cr = false; overflow = false;
  }
  __device__ __forceinline__ void raw_put(uint32_t c) { if (o < cap) out[o++] = (uint8_t)c; else overflow = true; }
  __device__ __forceinline__ void put(uint32_t c) {
    if (c == ' ' || c == '\t') { if (!prev_sp) { raw_put(' '); prev_sp = true; } return; }
    prev_sp = false;
    raw_put(c);
  }
  __device__ inline void feed(FeRd& R, uint32_t off, uint32_t n) {
    uint32_t i = 0;
    if (pending_cr && n) { pending_cr = false; if (R(off) == '\n') i = 1; else put('\r'); }
    while (i < n) {
      const uint32_t c = R(off + i);
      if (c == '\r') {
        if (i + 1 < n) { if (R(off + i + 1) == '\n') { i += 2; continue; } }
        else { pending_cr = true; i++; continue; }
      }
      put(c);
      i++;
    }
  }
  __device__ __forceinline__ uint32_t finish() {
    if (pending_cr) { pending_cr = false; put('\r'); }
    if (o > start && out[o - 1] == ' ') o--;
    raw_put('\r'); raw_put('\n');
    return o;
  }
};

__device__ __forceinline__ int fe_b64(uint32_t c) {
  if (c - 'A' < 26u) return (int)(c - 'A');
  if (c - 'a' < 26u) return (int)(c - 'a' + 26);
  if (c - '0' < 10u) return (int)(c - '0' + 52);
  if (c == '+') return 62;
  if (c == '/') return 63;
  return -1;
}

// Strict STANDARD base64 of a tag value with its FWS skipped.  Writes the decoded bytes through `sink(i, byte)`
// (i = index from the start).  Returns the decoded length or -1 on a syntax error.
template <typename Sink>
__device__ inline int fe_b64_decode(FeRd& R, uint32_t so, FeVal v, Sink sink) {
  uint32_t nchars = 0;
  for (uint32_t i = 0; i < v.len; i++) if (!fe_fws(R(so + v.off + i))) nchars++;
  if (nchars % 4) return -1;
  uint32_t q[4];
  uint32_t nq = 0, done = 0, o = 0;
  for (uint32_t i = 0; i < v.len; i++) {
    const uint32_t c = R(so + v.off + i);
    if (fe_fws(c)) continue;
    q[nq++] = c;
    if (nq < 4) continue;
    nq = 0;
    done += 4;
    const bool last = done == nchars;
    const int a = fe_b64(q[0]), b = fe_b64(q[1]), cc = fe_b64(q[2]), d = fe_b64(q[3]);
    if (a < 0 || b < 0) return -1;
    if (cc >= 0 && d >= 0) {
      sink(o++, (uint32_t)((a << 2) | (b >> 4)) & 0xffu);
      sink(o++, (uint32_t)((b << 4) | (cc >> 2)) & 0xffu);
      sink(o++, (uint32_t)((cc << 6) | d) & 0xffu);
      continue;
    }
    if (!last) return -1;
    if (q[2] == '=' && q[3] == '=') {
      if (b & 15) return -1;
      sink(o++, (uint32_t)((a << 2) | (b >> 4)) & 0xffu);
    } else if (q[3] == '=' && cc >= 0) {
      if (cc & 3) return -1;
      sink(o++, (uint32_t)((a << 2) | (b >> 4)) & 0xffu);
      sink(o++, (uint32_t)((b << 4) | (cc >> 2)) & 0xffu);
    } else return -1;
  }
  return (int)o;
}

// One message.  pre: the preimage slot (FE_PRE_CAP bytes); sigw: `limbs` words, zeroed here.
__device__ inline void fe_process(const uint8_t* raw, uint32_t n, const uint8_t* dom, uint32_t dom_len, uint32_t k,
                                  uint32_t limbs, uint8_t* pre, uint32_t* sigw, FeOut& out, uint32_t& body_l,
                                  bool allow_skip = false, long long now = 0) {
  body_l = 0;
  out.flags = 0; out.body_off = 0; out.body_len = 0; out.pre_len = 0;
  for (int i = 0; i < 8; i++) out.bh[i] = 0;
  for (uint32_t i = 0; i < limbs; i++) sigw[i] = 0;
  FeRd R;
  R.init(raw);
  FeHdr hs[FE_MAXH];
  uint32_t nh = 0, body_off = 0;
  const int pr = fe_parse_headers(R, n, hs, nh, body_off);
  if (pr == 1) { out.flags = FE_MAIL_PARSE; return; }
  if (pr == 2) { out.flags = FE_FALLBACK; return; }
  // body = bytes after the first CRLF CRLF, which is the end of the header block when the block ends that way
  if (!(body_off >= 4 && R(body_off - 4) == '\r' && R(body_off - 3) == '\n' && R(body_off - 2) == '\r' && R(body_off - 1) == '\n')) {
    out.flags = FE_FALLBACK; return;
  }
  out.body_off = body_off; out.body_len = n - body_off;
  // DKIM-Signature headers, top to bottom.  The reference walks them in order, skips those whose d= is another
  // domain and lets the first one that verifies win (cfdkim::verify_email_with_key).  The device takes the first
  // header of from_domain as its candidate, provided every header before it parses cleanly as a signature of another
  // domain; a message with several signature headers is marked FE_MULTI so that a candidate that does NOT verify
  // sends the message to the host front end (which tries the later headers and reports the reference's detail).
  // With regex parts the haystacks come from the first VALID header whatever its domain, so skipping is off.
  int sigs[8];
  uint32_t n_sigs = 0;
  for (uint32_t i = 0; i < nh; i++) {
    const FeHdr& h = hs[i];
    bool is_sig = h.key_len == 14;
    const char* lit = "dkim-signature";
    for (uint32_t j = 0; is_sig && j < 14; j++) is_sig = fe_lower(R(h.key_off + j)) == (uint32_t)(uint8_t)lit[j];
    if (is_sig) { if (n_sigs >= 8) { out.flags = FE_FALLBACK; return; } sigs[n_sigs++] = (int)i; }
  }
  if (n_sigs == 0) { out.flags = FE_FALLBACK; return; }
  uint32_t so = 0, sn = 0;   // the candidate's header value is R(so .. so + sn)
  FeVal tv, ta, tb, tbh, td, th, tc, ti, tq, tx, tl;
  uint32_t seen = 0;  // bit per slot: v1 a2 b4 bh8 d16 h32 c64 s128 i256 q512 x1024 l2048
  // 0: a well-formed rsa-sha256 signature of from_domain; 2: well-formed, another domain; 1: anything else
  auto parse_sig = [&](int idx) -> int {
    so = hs[idx].val_off;
    sn = hs[idx].val_len;
    for (uint32_t i = 0; i < sn; i++) if (R(so + i) & 0x80) return 1;
    // ---- tag list (cfdkim parser.rs grammar): slots v a b bh d h c s; i q x l and duplicates fall back
    tv.len = ta.len = tb.len = tbh.len = td.len = th.len = tc.len = ti.len = tq.len = tx.len = tl.len = 0;
    tv.off = ta.off = tb.off = tbh.off = td.off = th.off = tc.off = ti.off = tq.off = tx.off = tl.off = 0;
    seen = 0;
    uint32_t pos = 0;
    bool first = true;
    for (;;) {
      uint32_t p = pos;
      if (!first) { if (p >= sn || R(so + p) != ';') break; p++; }
      while (p < sn && fe_fws(R(so + p))) p++;
      if (p >= sn || !fe_alpha(R(so + p))) { if (first) return 1; break; }
      const uint32_t name_off = p;
      while (p < sn && fe_alnum_(R(so + p))) p++;
      const uint32_t name_len = p - name_off;
      while (p < sn && fe_fws(R(so + p))) p++;
      if (p >= sn || R(so + p) != '=') { if (first) return 1; break; }
      p++;
      while (p < sn && fe_fws(R(so + p))) p++;
      FeVal val; val.off = p; val.len = 0;
      if (p < sn && fe_valchar(R(so + p))) {
        for (;;) {
          while (p < sn && fe_valchar(R(so + p))) p++;
          val.len = p - val.off;
          uint32_t q = p;
          while (q < sn && fe_fws(R(so + q))) q++;
          if (q == p || q >= sn || !fe_valchar(R(so + q))) break;
          p = q;
        }
      }
      while (p < sn && fe_fws(R(so + p))) p++;
      uint32_t bit = 0;
      const uint32_t c0 = R(so + name_off), c1 = name_len > 1 ? R(so + name_off + 1) : 0;
      if (name_len == 1) {
        switch (c0) {
          case 'v': bit = 1; tv = val; break;
          case 'a': bit = 2; ta = val; break;
          case 'b': bit = 4; tb = val; break;
          case 'd': bit = 16; td = val; break;
          case 'h': bit = 32; th = val; break;
          case 'c': bit = 64; tc = val; break;
          case 's': bit = 128; break;
          case 'i': bit = 256; ti = val; break;
          case 'q': bit = 512; tq = val; break;
          case 'x': bit = 1024; tx = val; break;
          case 'l': bit = 2048; tl = val; break;
          default: break;
        }
      } else if (name_len == 2 && c0 == 'b' && c1 == 'h') { bit = 8; tbh = val; }
      if (bit) { if (seen & bit) return 1; seen |= bit; }
      // unknown tag names may repeat in the reference's map without changing what it reads; names that
      // collide with each other are irrelevant to verification, so they are ignored here
      pos = p;
      first = false;
    }
    // text the parser stopped at (a trailing ';', garbage) is ignored, as cfdkim's tag_list does
    if ((seen & (1 | 2 | 4 | 8 | 16 | 32 | 128)) != (1 | 2 | 4 | 8 | 16 | 32 | 128)) return 1;
    if (!fe_val_is(R, so, tv, "1") || !fe_val_is(R, so, ta, "rsa-sha256") || tb.len == 0) return 1;
    // optional tags the validator looks at (cfdkim validate_header): accepted when they pass, host path when they
    // do not (it reports the reference's error).  i= must end with the d= value (bytes, FWS removed); q= must be
    // dns/txt; x= must be a plain decimal clearly in the future of `now` (a signature near its expiry is left to the
    // host's own clock).
    if (seen & 256) {
      uint32_t li = 0, ld = 0;
      for (uint32_t i = 0; i < ti.len; i++) if (!fe_fws(R(so + ti.off + i))) li++;
      for (uint32_t i = 0; i < td.len; i++) if (!fe_fws(R(so + td.off + i))) ld++;
      if (li < ld) return 1;
      uint32_t a = ti.len, b = td.len;
      for (uint32_t m = 0; m < ld; m++) {
        do { a--; } while (fe_fws(R(so + ti.off + a)));
        do { b--; } while (fe_fws(R(so + td.off + b)));
        if (R(so + ti.off + a) != R(so + td.off + b)) return 1;
      }
    }
    if ((seen & 512) && !fe_val_is(R, so, tq, "dns/txt")) return 1;
    if (seen & 1024) {
      long long x = 0;
      uint32_t digits = 0;
      for (uint32_t i = 0; i < tx.len; i++) {
        const uint32_t c = R(so + tx.off + i);
        if (fe_fws(c)) continue;
        if (c < '0' || c > '9' || ++digits > 17) return 1;
        x = x * 10 + (long long)(c - '0');
      }
      if (digits == 0 || now + 2 > x + 15 * 60) return 1;
    }
    if (seen & 2048) {   // l=: [+]digits (cfdkim parses a usize); values beyond 32 bits never truncate a body that fits
      uint32_t digits = 0, lead = 0;
      unsigned long long v = 0;
      for (uint32_t i = 0; i < tl.len; i++) {
        const uint32_t c = R(so + tl.off + i);
        if (fe_fws(c)) continue;
        if (c == '+' && digits == 0 && lead == 0) { lead = 1; continue; }
        if (c < '0' || c > '9' || ++digits > 18) return 1;
        v = v * 10 + (unsigned long long)(c - '0');
      }
      if (digits == 0) return 1;
      body_l = v > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)v;
    }
    // d= == from_domain (ASCII case-insensitive, FWS removed)
    {
      uint32_t j = 0;
      bool ok = true;
      for (uint32_t i = 0; i < td.len && ok; i++) {
        const uint32_t c = R(so + td.off + i);
        if (fe_fws(c)) continue;
        ok = j < dom_len && fe_lower(c) == fe_lower(dom[j]);
        j++;
      }
      if (!ok || j != dom_len) return 2;   // a signature of another domain: the reference skips it
    }
    return 0;
  };
  int sig_idx = -1;
  for (uint32_t q = 0; q < n_sigs && sig_idx < 0; q++) {
    const int r = parse_sig(sigs[q]);
    if (r == 1 || (r == 2 && !allow_skip)) { out.flags = FE_FALLBACK; return; }
    if (r == 0) sig_idx = sigs[q];
  }
  if (sig_idx < 0) { out.flags = FE_FALLBACK; return; }
  const uint32_t multi = (n_sigs > 1 ? FE_MULTI : 0u) | ((seen & 2048) ? FE_HAS_L : 0u);
  bool hr = false, br = false;
  if (seen & 64) {
    if (fe_val_is(R, so, tc, "relaxed/relaxed")) { hr = true; br = true; }
    else if (fe_val_is(R, so, tc, "simple/simple") || fe_val_is(R, so, tc, "simple")) { hr = false; br = false; }
    else if (fe_val_is(R, so, tc, "relaxed/simple") || fe_val_is(R, so, tc, "relaxed")) { hr = true; br = false; }
    else if (fe_val_is(R, so, tc, "simple/relaxed")) { hr = false; br = true; }
    else { out.flags = FE_FALLBACK; return; }
  }
  // ---- h=: names (FWS removed) split on ':'; bottom-up selection with a per-name cursor
  uint32_t name_s[FE_MAXN], name_e[FE_MAXN];   // slices of the h value in "compact" coordinates
  uint8_t hbuf[256];                            // the h value with FWS removed
  uint32_t hl = 0;
  for (uint32_t i = 0; i < th.len; i++) {
    const uint32_t c = R(so + th.off + i);
    if (fe_fws(c)) continue;
    if (hl >= sizeof hbuf) { out.flags = FE_FALLBACK; return; }
    hbuf[hl++] = (uint8_t)c;
  }
  uint32_t nn = 0;
  bool has_from = false;
  {
    uint32_t a = 0;
    for (uint32_t i = 0; i <= hl; i++) {
      if (i != hl && hbuf[i] != ':') continue;
      if (i > a) {
        if (nn >= FE_MAXN) { out.flags = FE_FALLBACK; return; }
        name_s[nn] = a; name_e[nn] = i; nn++;
        if (i - a == 4 && fe_lower(hbuf[a]) == 'f' && fe_lower(hbuf[a + 1]) == 'r' && fe_lower(hbuf[a + 2]) == 'o' && fe_lower(hbuf[a + 3]) == 'm') has_from = true;
      }
      a = i + 1;
    }
  }
  if (!has_from) { out.flags = FE_FALLBACK; return; }
  int hit_of[FE_MAXN];
  uint32_t o = 0;
  bool overflow = false;
  for (uint32_t j = 0; j < nn; j++) {
    const uint32_t nl = name_e[j] - name_s[j];
    int start = (int)nh;
    for (uint32_t i = 0; i < j; i++) {  // the latest earlier occurrence of the same name sets the cursor
      if (name_e[i] - name_s[i] != nl) continue;
      bool same = true;
      for (uint32_t t = 0; same && t < nl; t++) same = fe_lower(hbuf[name_s[i] + t]) == fe_lower(hbuf[name_s[j] + t]);
      if (same) start = hit_of[i] >= 0 ? hit_of[i] : 0;
    }
    int hit = -1;
    for (int x = start - 1; x >= 0; x--) {
      const FeHdr& h = hs[x];
      if (h.key_len != nl) continue;
      bool same = true;
      for (uint32_t t = 0; same && t < nl; t++) same = fe_lower(R(h.key_off + t)) == fe_lower(hbuf[name_s[j] + t]);
      if (same) { hit = x; break; }
    }
    hit_of[j] = hit;
    if (hit < 0) continue;
    const FeHdr& h = hs[hit];
    for (uint32_t t = 0; t < h.key_len; t++) if (R(h.key_off + t) & 0x80) { out.flags = FE_FALLBACK; return; }
    if (hr) {
      uint32_t kl = h.key_len;
      while (kl > 0 && (R(h.key_off + kl - 1) == ' ' || (R(h.key_off + kl - 1) >= 9 && R(h.key_off + kl - 1) <= 13))) kl--;
      for (uint32_t t = 0; t < kl; t++) { if (o < FE_PRE_CAP) pre[o++] = (uint8_t)fe_lower(R(h.key_off + t)); else overflow = true; }
      if (o < FE_PRE_CAP) pre[o++] = ':'; else overflow = true;
      FeRelaxed rv;
      rv.init(pre, o, FE_PRE_CAP);
      rv.feed(R, h.val_off, h.val_len);
      o = rv.finish();
      overflow = overflow || rv.overflow;
    } else {
      for (uint32_t t = 0; t < h.key_len; t++) { if (o < FE_PRE_CAP) pre[o++] = R(h.key_off + t); else overflow = true; }
      if (o + 2 <= FE_PRE_CAP) { pre[o++] = ':'; pre[o++] = ' '; } else overflow = true;
      for (uint32_t t = 0; t < h.val_len; t++) { if (o < FE_PRE_CAP) pre[o++] = R(h.val_off + t); else overflow = true; }
      if (o + 2 <= FE_PRE_CAP) { pre[o++] = '\r'; pre[o++] = '\n'; } else overflow = true;
    }
  }
  // ---- the signature header with the raw b= text removed (value.replace(raw_b, "")); another occurrence of
  // that text anywhere in the value is left to the host
  for (uint32_t p = 0; p + tb.len <= sn; p++) {
    if (p == tb.off || R(so + p) != R(so + tb.off)) continue;
    bool same = true;
    for (uint32_t t = 1; same && t < tb.len; t++) same = R(so + p + t) == R(so + tb.off + t);
    if (same) { out.flags = FE_FALLBACK; return; }
  }
  if (hr) {
    const char* kn = "dkim-signature:";
    for (uint32_t t = 0; t < 15; t++) { if (o < FE_PRE_CAP) pre[o++] = (uint8_t)kn[t]; else overflow = true; }
    FeRelaxed rv;
    rv.init(pre, o, FE_PRE_CAP);
    rv.feed(R, so, tb.off);
    rv.feed(R, so + tb.off + tb.len, sn - tb.off - tb.len);
    o = rv.finish();
    overflow = overflow || rv.overflow;
  } else {
    const char* kn = "DKIM-Signature: ";
    for (uint32_t t = 0; t < 16; t++) { if (o < FE_PRE_CAP) pre[o++] = (uint8_t)kn[t]; else overflow = true; }
    for (uint32_t t = 0; t < sn; t++) {
      if (t >= tb.off && t < tb.off + tb.len) continue;
      if (o < FE_PRE_CAP) pre[o++] = R(so + t); else overflow = true;
    }
    if (o + 2 <= FE_PRE_CAP) { pre[o++] = '\r'; pre[o++] = '\n'; } else overflow = true;
  }
  if (overflow || o < 2) { out.flags = FE_FALLBACK; return; }
  out.pre_len = o - 2;  // final CRLF dropped
  uint32_t flags = (hr ? FE_HDR_RELAXED : 0u) | (br ? FE_BODY_RELAXED : 0u) | multi;
  // ---- bh= : 44 base64 characters -> 32 bytes -> 8 big-endian words
  {
    uint8_t bhb[48];
    uint32_t cnt = 0;
    for (uint32_t i = 0; i < tbh.len; i++) if (!fe_fws(R(so + tbh.off + i))) cnt++;
    int dl = -1;
    if (cnt == 44) dl = fe_b64_decode(R, so, tbh, [&](uint32_t i, uint32_t b) { if (i < 48) bhb[i] = (uint8_t)b; });
    if (dl == 32) {
      flags |= FE_BH_VALID;
      for (int i = 0; i < 8; i++)
        out.bh[i] = ((uint32_t)bhb[4 * i] << 24) | ((uint32_t)bhb[4 * i + 1] << 16) | ((uint32_t)bhb[4 * i + 2] << 8) | bhb[4 * i + 3];
    }
  }
  // ---- b= : signature bytes (big endian) -> little-endian limbs; the decoded length must equal k
  {
    // first pass: length only; second pass writes the limbs when the length is right
    const int sl = fe_b64_decode(R, so, tb, [](uint32_t, uint32_t) {});
    if (sl < 0) flags |= FE_SIG_SYNTAX;
    else if ((uint32_t)sl != k || k > 4 * limbs) flags |= FE_SIG_BADLEN;
    else fe_b64_decode(R, so, tb, [&](uint32_t i, uint32_t b) {
      const uint32_t bi = (uint32_t)sl - 1 - i;
      sigw[bi >> 2] |= b << (8 * (bi & 3));
    });
  }
  out.flags = flags;
}

// One thread per message.  Also emits the body's CanonItem (an empty one for fallback / error messages).
__global__ void __launch_bounds__(128)
frontend_kernel(const uint8_t* __restrict__ span, const FeIn* __restrict__ in, uint32_t n, uint8_t* __restrict__ arena,
                const uint64_t* __restrict__ msg_off, uint32_t* __restrict__ msg_len, uint32_t* __restrict__ sig_arena,
                uint32_t* __restrict__ cand_bh, CanonItem* __restrict__ canon, FeOut* __restrict__ out, int allow_skip,
                long long now) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const FeIn fi = in[idx];
  FeOut fo;
  uint32_t body_l = 0;
  fe_process(span + fi.raw_off, fi.raw_len, arena + fi.dom_off, fi.dom_len, fi.k, fi.limbs, arena + msg_off[fi.pre_msg],
             sig_arena + fi.sig_word_off, fo, body_l, allow_skip != 0, now);
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
