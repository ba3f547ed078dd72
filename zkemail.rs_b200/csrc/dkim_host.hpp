// dkim_host.hpp — host front end of the engine: RFC 5322 header split (mailparse 0.15 rules),
// DKIM-Signature tag-list parsing + validation, signed-header selection, simple/relaxed header and
// body canonicalisation, base64.  This is the part of cfdkim::verify_email_with_key /
// canonicalize_signed_email (core/src/email.rs:31-33, core/src/circuits.rs:34-35) that is byte
// shuffling rather than arithmetic; everything it emits goes to the device kernels.
// Semantics: SURVEY.md Appendix A.2.  Independent single-pass implementation (the oracle in
// oracle/zk_oracle.c follows the reference's multi-pass operation order instead).
#pragma once
#include <stdint.h>
#include <string.h>
#include <time.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define ZKB_HAVE_AVX2_DISPATCH 1
#endif

#include <string>
#include <vector>

#include "../../include/zkemail_b200.h"

namespace zkb {

struct HeaderField {
  uint32_t key_off, key_len, val_off, val_len;
};

// mailparse::parse_headers.  false => parse_mail() would return Err (reference panics).
inline bool parse_headers(const uint8_t* d, size_t n, std::vector<HeaderField>& out, size_t& body_off) {
  out.clear();
  size_t ix = 0;
  while (ix < n) {
    uint8_t c0 = d[ix];
    if (c0 == '\n') { ix++; break; }
    if (c0 == '\r') {
      if (ix + 1 < n && d[ix + 1] == '\n') { ix += 2; break; }
      return false;  // lone CR after the headers
    }
    if (c0 == ' ') return false;  // overhanging continuation line
    // key: up to ':' ; a '\n' before any ':' makes a value-less header ending at that line
    size_t p = ix;
    while (p < n && d[p] != ':' && d[p] != '\n') p++;
    HeaderField h;
    h.key_off = (uint32_t)ix;
    if (p >= n) {  // ran out of input inside the key: empty key, empty value
      h.key_len = 0; h.val_off = (uint32_t)ix; h.val_len = 0;
      out.push_back(h);
      ix = n;
      break;
    }
    if (d[p] == '\n') {
      h.key_len = (uint32_t)(p - ix); h.val_off = (uint32_t)p; h.val_len = 0;
      out.push_back(h);
      ix = p + 1;
      continue;
    }
    h.key_len = (uint32_t)(p - ix);
    p++;  // past ':'
    while (p < n && d[p] == ' ') p++;
    size_t vs = p, ve = p;
    // value: until a '\n' not followed by SP/TAB; end = one past the last byte that is not CR/LF
    for (;;) {
      const uint8_t* nl = (const uint8_t*)memchr(d + p, '\n', n - p);
      size_t q = nl ? (size_t)(nl - d) : n;
      // last non-CR byte in [p,q)
      size_t e = q;
      while (e > p && d[e - 1] == '\r') e--;
      if (e > p) ve = e;
      if (!nl) { p = n; break; }
      p = q + 1;
      if (p < n && (d[p] == ' ' || d[p] == '\t')) continue;
      break;
    }
    if (ve < vs) ve = vs;
    h.val_off = (uint32_t)vs; h.val_len = (uint32_t)(ve - vs);
    out.push_back(h);
    ix = p;
  }
  body_off = ix;
  return true;
}

inline bool ieq_ascii(const uint8_t* a, size_t al, const char* b, size_t bl) {
  if (al != bl) return false;
  for (size_t i = 0; i < al; i++) {
    uint8_t x = a[i], y = (uint8_t)b[i];
    if ((unsigned)(x - 'A') < 26u) x += 32;
    if ((unsigned)(y - 'A') < 26u) y += 32;
    if (x != y) return false;
  }
  return true;
}

// String::from_utf8_lossy (only called when a byte >= 0x80 is present)
inline void utf8_lossy(const uint8_t* in, size_t n, std::string& out) {
  out.clear();
  size_t i = 0;
  auto at = [&](size_t k) -> uint8_t { return k < n ? in[k] : 0; };
  while (i < n) {
    uint8_t b = in[i];
    size_t w = 0, bad = 1;
    if (b < 0x80) w = 1;
    else if (b >= 0xC2 && b <= 0xDF) { if ((at(i + 1) & 0xC0) == 0x80) w = 2; }
    else if (b >= 0xE0 && b <= 0xEF) {
      uint8_t c = at(i + 1);
      bool ok2 = b == 0xE0 ? (c >= 0xA0 && c <= 0xBF) : b == 0xED ? (c >= 0x80 && c <= 0x9F) : (c >= 0x80 && c <= 0xBF);
      if (ok2) { if ((at(i + 2) & 0xC0) == 0x80) w = 3; else bad = 2; }
    } else if (b >= 0xF0 && b <= 0xF4) {
      uint8_t c = at(i + 1);
      bool ok2 = b == 0xF0 ? (c >= 0x90 && c <= 0xBF) : b == 0xF4 ? (c >= 0x80 && c <= 0x8F) : (c >= 0x80 && c <= 0xBF);
      if (ok2) {
        if ((at(i + 2) & 0xC0) != 0x80) bad = 2;
        else if ((at(i + 3) & 0xC0) != 0x80) bad = 3;
        else w = 4;
      }
    }
    if (w) { out.append((const char*)in + i, w); i += w; }
    else { out.append("\xEF\xBF\xBD", 3); i += bad; }
  }
}

// ------------------------------------------------------------------ tag list
struct Tag {
  uint32_t name_off, name_len;
  uint32_t raw_off, raw_len;   // value text with inner FWS kept (slice of the header value)
  uint32_t val_off, val_len;   // value with FWS removed (slice of DkimSig::vals)
};
struct DkimSig {
  const uint8_t* s = nullptr;  // header value (lossy utf-8 if it had non-ASCII bytes)
  size_t n = 0;
  std::string lossy, vals;
  std::vector<Tag> tags;
  // O(1) lookup of the tag names the verifier asks for: v a b d h s c l i q x (by letter) and bh
  int16_t known[27];
  static int slot_of(const uint8_t* name, size_t len) {
    if (len == 1 && name[0] >= 'a' && name[0] <= 'z') return name[0] - 'a';
    if (len == 2 && name[0] == 'b' && name[1] == 'h') return 26;
    return -1;
  }
  void reset() { tags.clear(); vals.clear(); for (auto& k : known) k = -1; }
  const Tag* get(const char* name) const {
    size_t l = strlen(name);
    int k = slot_of((const uint8_t*)name, l);
    if (k >= 0) return known[k] >= 0 ? &tags[known[k]] : nullptr;
    for (const Tag& t : tags)
      if (t.name_len == l && memcmp(s + t.name_off, name, l) == 0) return &t;
    return nullptr;
  }
  const uint8_t* val(const Tag* t) const { return (const uint8_t*)vals.data() + t->val_off; }
  bool val_is(const Tag* t, const char* lit) const {
    size_t l = strlen(lit);
    return t->val_len == l && memcmp(val(t), lit, l) == 0;
  }
};

namespace detail {
enum { C_FWS = 1, C_VAL = 2, C_ALPHA = 4, C_ALNUM = 8 };
struct CharTab {
  uint8_t t[256];
  CharTab() {
    for (int c = 0; c < 256; c++) {
      uint8_t v = 0;
      if (c == ' ' || c == '\t' || c == '\r' || c == '\n') v |= C_FWS;
      if ((c >= 0x21 && c <= 0x3A) || (c >= 0x3C && c <= 0x7E)) v |= C_VAL;
      bool al = (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z');
      if (al) v |= C_ALPHA;
      if (al || (c >= '0' && c <= '9') || c == '_') v |= C_ALNUM;
      t[c] = v;
    }
  }
};
inline const uint8_t* chartab() { static const CharTab T; return T.t; }
}  // namespace detail

// cfdkim parser::tag_list + lib.rs::validate_header.  Returns ZKB_DKIM_PASS or the error kind.
inline int validate_dkim_header(const uint8_t* raw_val, size_t raw_len, int64_t now_unix, DkimSig& sig) {
  using namespace detail;
  const uint8_t* CT = chartab();
  sig.reset();
  bool ascii = true;
  {
    size_t i = 0;
    uint64_t acc = 0;
    for (; i + 8 <= raw_len; i += 8) { uint64_t w; memcpy(&w, raw_val + i, 8); acc |= w; }
    for (; i < raw_len; i++) acc |= raw_val[i];
    ascii = (acc & 0x8080808080808080ull) == 0;
  }
  if (ascii) { sig.s = raw_val; sig.n = raw_len; }
  else { utf8_lossy(raw_val, raw_len, sig.lossy); sig.s = (const uint8_t*)sig.lossy.data(); sig.n = sig.lossy.size(); }
  const uint8_t* s = sig.s;
  const size_t n = sig.n;
  sig.vals.reserve(n);
  size_t pos = 0;
  bool first = true;
  for (;;) {
    size_t p = pos;
    if (!first) {
      if (p >= n || s[p] != ';') break;
      p++;
    }
    while (p < n && (CT[s[p]] & C_FWS)) p++;
    if (p >= n || !(CT[s[p]] & C_ALPHA)) { if (first) return ZKB_DKIM_SYNTAX; break; }
    Tag t;
    t.name_off = (uint32_t)p;
    while (p < n && (CT[s[p]] & C_ALNUM)) p++;
    t.name_len = (uint32_t)(p - t.name_off);
    while (p < n && (CT[s[p]] & C_FWS)) p++;
    if (p >= n || s[p] != '=') { if (first) return ZKB_DKIM_SYNTAX; break; }
    p++;
    while (p < n && (CT[s[p]] & C_FWS)) p++;
    t.raw_off = (uint32_t)p; t.raw_len = 0;
    t.val_off = (uint32_t)sig.vals.size();
    if (p < n && (CT[s[p]] & C_VAL)) {
      for (;;) {
        size_t a = p;
        while (p < n && (CT[s[p]] & C_VAL)) p++;
        sig.vals.append((const char*)s + a, p - a);
        t.raw_len = (uint32_t)(p - t.raw_off);
        size_t q = p;
        while (q < n && (CT[s[q]] & C_FWS)) q++;
        if (q == p || q >= n || !(CT[s[q]] & C_VAL)) break;
        p = q;
      }
    }
    t.val_len = (uint32_t)(sig.vals.size() - t.val_off);
    while (p < n && (CT[s[p]] & C_FWS)) p++;
    // IndexMap insert: a later duplicate replaces the value in the first one's position
    const int slot = DkimSig::slot_of(s + t.name_off, t.name_len);
    bool dup = false;
    if (slot >= 0) {
      if (sig.known[slot] >= 0) { sig.tags[sig.known[slot]] = t; dup = true; }
    } else {
      for (Tag& o : sig.tags)
        if (o.name_len == t.name_len && memcmp(s + o.name_off, s + t.name_off, t.name_len) == 0) { o = t; dup = true; break; }
    }
    if (!dup) {
      if (slot >= 0 && sig.tags.size() < 32000) sig.known[slot] = (int16_t)sig.tags.size();
      sig.tags.push_back(t);
    }
    pos = p;
    first = false;
  }
  static const char* REQ[] = {"v", "a", "b", "bh", "d", "h", "s"};
  for (const char* r : REQ) if (!sig.get(r)) return ZKB_DKIM_MISSING_TAG;
  if (!sig.val_is(sig.get("v"), "1")) return ZKB_DKIM_VERSION;
  const Tag* td = sig.get("d");
  if (const Tag* ti = sig.get("i")) {
    if (ti->val_len < td->val_len || memcmp(sig.val(ti) + ti->val_len - td->val_len, sig.val(td), td->val_len) != 0)
      return ZKB_DKIM_DOMAIN_MISMATCH;
  }
  {
    const Tag* th = sig.get("h");
    const uint8_t* v = sig.val(th);
    bool found = false;
    size_t a = 0;
    for (size_t i = 0; i <= th->val_len; i++)
      if (i == th->val_len || v[i] == ':') { if (ieq_ascii(v + a, i - a, "from", 4)) found = true; a = i + 1; }
    if (!found) return ZKB_DKIM_FROM_NOT_SIGNED;
  }
  if (const Tag* tq = sig.get("q")) if (!sig.val_is(tq, "dns/txt")) return ZKB_DKIM_QUERY_METHOD;
  if (const Tag* tx = sig.get("x")) {
    const uint8_t* v = sig.val(tx);
    size_t l = tx->val_len, i = 0;
    bool ok = l > 0, neg = false;
    int64_t x = 0;
    if (ok && (v[0] == '+' || v[0] == '-')) { neg = v[0] == '-'; i = 1; ok = l > 1; }
    for (; ok && i < l; i++) {
      if (v[i] < '0' || v[i] > '9') { ok = false; break; }
      int dgt = v[i] - '0';
      if (!neg) { if (x > (INT64_MAX - dgt) / 10) { ok = false; break; } x = x * 10 + dgt; }
      else { if (x < (INT64_MIN + dgt) / 10) { ok = false; break; } x = x * 10 - dgt; }
    }
    if (!ok) x = 0;
    int64_t now = now_unix ? now_unix : (int64_t)time(nullptr);
    if (now > x + 15 * 60) return ZKB_DKIM_EXPIRED;
  }
  return ZKB_DKIM_PASS;
}

// c= tag.  false => UnsupportedCanonicalizationType
inline bool parse_canon_tag(const DkimSig& sig, bool& hdr_relaxed, bool& body_relaxed) {
  const Tag* tc = sig.get("c");
  if (!tc) { hdr_relaxed = false; body_relaxed = false; return true; }
  if (sig.val_is(tc, "relaxed/relaxed")) { hdr_relaxed = true; body_relaxed = true; }
  else if (sig.val_is(tc, "simple/simple") || sig.val_is(tc, "simple")) { hdr_relaxed = false; body_relaxed = false; }
  else if (sig.val_is(tc, "relaxed/simple") || sig.val_is(tc, "relaxed")) { hdr_relaxed = true; body_relaxed = false; }
  else if (sig.val_is(tc, "simple/relaxed")) { hdr_relaxed = false; body_relaxed = true; }
  else return false;
  return true;
}

// usize::from_str for the l= tag
inline bool parse_usize_tag(const DkimSig& sig, const Tag* t, uint64_t& out) {
  const uint8_t* v = sig.val(t);
  size_t i = 0;
  if (t->val_len && v[0] == '+') i = 1;
  if (i >= t->val_len) return false;
  uint64_t x = 0;
  for (; i < t->val_len; i++) {
    if (v[i] < '0' || v[i] > '9') return false;
    if (x > (UINT64_MAX - (v[i] - '0')) / 10) return false;
    x = x * 10 + (v[i] - '0');
  }
  out = x;
  return true;
}

// ------------------------------------------------------------------ canonicalisation (single pass)
// Relaxed body; out must hold n + 2 bytes (+32 bytes of slack for vector stores).  Single pass with
// the state (o, prev_sp); chunks that the state machine would copy verbatim (no TAB, no SP followed
// by SP/CR, not ending in SP, not starting with LF, and not directly after a SP) take a vector fast
// path: 32 bytes per step with AVX2 (runtime dispatch), else 16 with SSE2.
struct RelaxedBody {
  const uint8_t* in;
  uint8_t* out;
  size_t n, o = 0, i = 0;
  bool prev_sp = false;
  inline void scalar(size_t end) {
    for (; i < end; i++) {
      uint8_t c = in[i];
      if (c == ' ' || c == '\t') {
        if (!prev_sp) { out[o++] = ' '; prev_sp = true; }
        continue;
      }
      prev_sp = false;
      if (c == '\n' && o >= 2 && out[o - 1] == '\r' && out[o - 2] == ' ') {
        out[o - 2] = '\r'; out[o - 1] = '\n';  // drop the single SP before CRLF
        continue;
      }
      out[o++] = c;
    }
  }
  inline size_t finish() {
    scalar(n);
    while (o >= 4 && out[o - 1] == '\n' && out[o - 2] == '\r' && out[o - 3] == '\n' && out[o - 4] == '\r') o -= 2;
    if (o > 0 && !(o >= 2 && out[o - 2] == '\r' && out[o - 1] == '\n')) { out[o++] = '\r'; out[o++] = '\n'; }
    return o;
  }
};
#if defined(ZKB_HAVE_AVX2_DISPATCH)
__attribute__((target("avx2"))) inline void relaxed_body_avx2(RelaxedBody& st) {
  const __m256i vsp = _mm256_set1_epi8(' '), vtab = _mm256_set1_epi8('\t'), vcr = _mm256_set1_epi8('\r');
  while (st.i + 32 <= st.n) {
    __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(st.in + st.i));
    uint32_t sp = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, vsp));
    uint32_t tab = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, vtab));
    uint32_t cr = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, vcr));
    uint32_t bad = tab | (sp & ((sp | cr) >> 1)) | (sp & 0x80000000u);
    if (bad | (uint32_t)st.prev_sp | (uint32_t)(st.in[st.i] == '\n')) { st.scalar(st.i + 32); continue; }
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(st.out + st.o), v);
    st.o += 32; st.i += 32;
  }
}
inline bool cpu_has_avx2() { static const bool v = __builtin_cpu_supports("avx2"); return v; }
#endif
inline size_t canon_body_relaxed(const uint8_t* in, size_t n, uint8_t* out) {
  RelaxedBody st;
  st.in = in; st.out = out; st.n = n;
#if defined(ZKB_HAVE_AVX2_DISPATCH)
  if (cpu_has_avx2()) { relaxed_body_avx2(st); return st.finish(); }
#endif
#if defined(__SSE2__)
  const __m128i vsp = _mm_set1_epi8(' '), vtab = _mm_set1_epi8('\t'), vcr = _mm_set1_epi8('\r');
  while (st.i + 16 <= n) {
    __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(in + st.i));
    unsigned sp = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(v, vsp));
    unsigned tab = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(v, vtab));
    unsigned cr = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(v, vcr));
    unsigned bad = tab | (sp & ((sp | cr) >> 1)) | (sp & 0x8000u);
    if (bad | (unsigned)st.prev_sp | (unsigned)(in[st.i] == '\n')) { st.scalar(st.i + 16); continue; }
    _mm_storeu_si128(reinterpret_cast<__m128i*>(out + st.o), v);
    st.o += 16; st.i += 16;
  }
#endif
  return st.finish();
}
inline size_t canon_body_simple(const uint8_t* in, size_t n, uint8_t* out) {
  if (n == 0) { out[0] = '\r'; out[1] = '\n'; return 2; }
  while (n >= 4 && in[n - 1] == '\n' && in[n - 2] == '\r' && in[n - 3] == '\n' && in[n - 4] == '\r') n -= 2;
  memcpy(out, in, n);
  return n;
}
inline bool latin1_ws(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13) || c == 0x85 || c == 0xA0; }
inline size_t put_key(const uint8_t* key, size_t klen, uint8_t* out, bool lower) {
  size_t o = 0;
  for (size_t i = 0; i < klen; i++) {
    uint32_t c = key[i];
    if (lower) {
      if (c - 'A' < 26u) c += 32;
      else if (c >= 0xC0 && c <= 0xDE && c != 0xD7) c += 32;
    }
    if (c < 0x80) out[o++] = (uint8_t)c;
    else { out[o++] = (uint8_t)(0xC0 | (c >> 6)); out[o++] = (uint8_t)(0x80 | (c & 0x3F)); }
  }
  return o;
}
// Relaxed header value, streamed over one or more input segments (the b= blanking of the
// DKIM-Signature header yields two): TAB->SP, every CRLF removed, SP runs collapsed, leading and
// trailing SP dropped.  A CR at the end of one segment and a LF at the start of the next form a
// CRLF, exactly as if the segments had been concatenated first.
struct RelaxedValue {
  uint8_t* out;
  size_t o, start;
  bool prev_sp = true;     // swallows leading SP
  bool pending_cr = false;
  explicit RelaxedValue(uint8_t* dst, size_t at) : out(dst), o(at), start(at) {}
  inline void put(uint8_t c) {
    if (c == ' ' || c == '\t') {
      if (!prev_sp) { out[o++] = ' '; prev_sp = true; }
      return;
    }
    prev_sp = false;
    out[o++] = c;
  }
  void feed(const uint8_t* v, size_t n) {
    size_t i = 0;
    if (pending_cr && n) {
      pending_cr = false;
      if (v[0] == '\n') i = 1; else put('\r');
    }
    while (i < n) {
#if defined(__SSE2__)
      if (i + 16 <= n && !prev_sp) {
        __m128i x = _mm_loadu_si128(reinterpret_cast<const __m128i*>(v + i));
        unsigned sp = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(x, _mm_set1_epi8(' ')));
        unsigned bad = (unsigned)_mm_movemask_epi8(_mm_or_si128(_mm_cmpeq_epi8(x, _mm_set1_epi8('\t')), _mm_cmpeq_epi8(x, _mm_set1_epi8('\r'))));
        if (!(bad | (sp & (sp >> 1)) | (sp & 0x8000u))) {
          _mm_storeu_si128(reinterpret_cast<__m128i*>(out + o), x);
          o += 16; i += 16;
          continue;
        }
      }
#endif
      uint8_t c = v[i];
      if (c == '\r') {
        if (i + 1 < n) { if (v[i + 1] == '\n') { i += 2; continue; } }
        else { pending_cr = true; i++; continue; }
      }
      put(c);
      i++;
    }
  }
  size_t finish() {
    if (pending_cr) { pending_cr = false; put('\r'); }
    if (o > start && out[o - 1] == ' ') o--;
    out[o++] = '\r'; out[o++] = '\n';
    return o;
  }
};
// out must hold 2*klen + vlen + 3 (+16 bytes of slack for the vector stores)
inline size_t canon_header_relaxed(const uint8_t* key, size_t klen, const uint8_t* val, size_t vlen, uint8_t* out) {
  while (klen > 0 && latin1_ws(key[klen - 1])) klen--;
  size_t o = put_key(key, klen, out, true);
  out[o++] = ':';
  RelaxedValue rv(out, o);
  rv.feed(val, vlen);
  return rv.finish();
}
inline size_t canon_header_simple(const uint8_t* key, size_t klen, const uint8_t* val, size_t vlen, uint8_t* out) {
  size_t o = put_key(key, klen, out, false);
  out[o++] = ':'; out[o++] = ' ';
  memcpy(out + o, val, vlen); o += vlen;
  out[o++] = '\r'; out[o++] = '\n';
  return o;
}

// Upper bound of the header-hash preimage for a message whose header block is `hdr_bytes` long.
inline size_t preimage_bound(size_t hdr_bytes, size_t sig_val_len) { return 2 * hdr_bytes + 3 * sig_val_len + 96; }

// select_headers + canonicalise + the b-less DKIM-Signature (no trailing CRLF). Returns length.
inline size_t build_header_preimage(const uint8_t* raw, const std::vector<HeaderField>& hs, const DkimSig& sig,
                                    bool relaxed, uint8_t* out, std::string& scratch) {
  const Tag* th = sig.get("h");
  const uint8_t* hv = sig.val(th);
  size_t o = 0;
  struct Cur { uint32_t off, len; long idx; };
  Cur last[64];
  int n_last = 0;
  std::vector<Cur> more;  // beyond 64 distinct names (pathological)
  size_t a = 0;
  for (size_t i = 0; i <= th->val_len; i++) {
    if (i != th->val_len && hv[i] != ':') continue;
    size_t s = a, e = i;
    a = i + 1;
    while (s < e && (hv[s] == ' ' || (hv[s] >= 9 && hv[s] <= 13))) s++;  // cannot occur (FWS stripped)
    while (e > s && (hv[e - 1] == ' ' || (hv[e - 1] >= 9 && hv[e - 1] <= 13))) e--;
    if (s == e) continue;
    long start = (long)hs.size();
    Cur* cur = nullptr;
    for (int k = 0; k < n_last; k++)
      if (ieq_ascii(hv + last[k].off, last[k].len, (const char*)hv + s, e - s)) { cur = &last[k]; break; }
    if (!cur) for (Cur& c : more) if (ieq_ascii(hv + c.off, c.len, (const char*)hv + s, e - s)) { cur = &c; break; }
    if (cur) start = cur->idx;
    long hit = -1;
    for (long j = start - 1; j >= 0; j--)
      if (ieq_ascii(raw + hs[j].key_off, hs[j].key_len, (const char*)hv + s, e - s)) { hit = j; break; }
    if (!cur) {
      if (n_last < 64) cur = &last[n_last++];
      else { more.push_back(Cur{}); cur = &more.back(); }
      cur->off = (uint32_t)s; cur->len = (uint32_t)(e - s);
    }
    cur->idx = hit >= 0 ? hit : 0;
    if (hit >= 0) {
      const HeaderField& h = hs[hit];
      o += relaxed ? canon_header_relaxed(raw + h.key_off, h.key_len, raw + h.val_off, h.val_len, out + o)
                   : canon_header_simple(raw + h.key_off, h.key_len, raw + h.val_off, h.val_len, out + o);
    }
  }
  // the signature header with every occurrence of the raw b= text removed
  // (value.replace(raw_b, ""): non-overlapping, left to right), canonicalised, final CRLF dropped
  const Tag* tb = sig.get("b");
  const uint8_t* v = sig.s;
  const size_t vl = sig.n;
  const uint8_t* pat = sig.s + tb->raw_off;
  const size_t pl = tb->raw_len;
  if (relaxed) {
    memcpy(out + o, "dkim-signature:", 15);
    RelaxedValue rv(out, o + 15);
    size_t i = 0, copied = 0;
    if (pl) {
      while (i + pl <= vl) {
        const uint8_t* f = (const uint8_t*)memchr(v + i, pat[0], vl - pl - i + 1);
        if (!f) break;
        i = (size_t)(f - v);
        if (memcmp(f, pat, pl) == 0) { rv.feed(v + copied, i - copied); i += pl; copied = i; }
        else i++;
      }
    }
    rv.feed(v + copied, vl - copied);
    return rv.finish() - 2;
  }
  scratch.clear();
  {
    size_t i = 0, copied = 0;
    if (pl) {
      while (i + pl <= vl) {
        const uint8_t* f = (const uint8_t*)memchr(v + i, pat[0], vl - pl - i + 1);
        if (!f) break;
        i = (size_t)(f - v);
        if (memcmp(f, pat, pl) == 0) { scratch.append((const char*)v + copied, i - copied); i += pl; copied = i; }
        else i++;
      }
    }
    scratch.append((const char*)v + copied, vl - copied);
  }
  size_t w = canon_header_simple((const uint8_t*)"DKIM-Signature", 14, (const uint8_t*)scratch.data(), scratch.size(), out + o);
  return o + w - 2;
}

// bytes::get_all_after(raw, "\r\n\r\n").  hdr_end (optional) = end of the header block found by
// parse_headers: when the block ends in CRLF CRLF that is necessarily the FIRST CRLF CRLF of the
// message (an earlier one would have ended the block earlier), so no search is needed.
inline const uint8_t* find_body(const uint8_t* raw, size_t n, size_t& blen, size_t hdr_end = 0) {
  if (hdr_end >= 4 && hdr_end <= n && memcmp(raw + hdr_end - 4, "\r\n\r\n", 4) == 0) {
    blen = n - hdr_end;
    return raw + hdr_end;
  }
  const uint8_t* f = n >= 4 ? (const uint8_t*)memmem(raw, n, "\r\n\r\n", 4) : nullptr;
  if (!f) { blen = 0; return raw + n; }
  blen = n - (size_t)(f - raw) - 4;
  return f + 4;
}

// ------------------------------------------------------------------ base64 (STANDARD, strict)
struct B64 {
  int8_t t[256];
  uint32_t d0[256], d1[256], d2[256], d3[256];  // pre-shifted values; 0x01000000 flags an invalid char
  B64() {
    memset(t, -1, sizeof t);
    const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    for (int i = 0; i < 64; i++) t[(uint8_t)a[i]] = (int8_t)i;
    for (int c = 0; c < 256; c++) {
      if (t[c] < 0) { d0[c] = d1[c] = d2[c] = d3[c] = 0x01000000u; continue; }
      uint32_t v = (uint32_t)t[c];
      d0[c] = v << 18; d1[c] = v << 12; d2[c] = v << 6; d3[c] = v;
    }
  }
};
inline const B64& b64tab() { static const B64 t; return t; }
// returns decoded length or -1.  out needs 3*n/4 bytes.
inline long base64_decode(const uint8_t* in, size_t n, uint8_t* out) {
  if (n % 4) return -1;
  const B64& B = b64tab();
  const int8_t* T = B.t;
  size_t o = 0, i = 0;
  for (; i + 4 < n; i += 4) {  // all quads but the last cannot carry padding
    uint32_t v = B.d0[in[i]] | B.d1[in[i + 1]] | B.d2[in[i + 2]] | B.d3[in[i + 3]];
    if (v & 0x01000000u) return -1;
    out[o++] = (uint8_t)(v >> 16); out[o++] = (uint8_t)(v >> 8); out[o++] = (uint8_t)v;
  }
  if (i < n) {
    int a = T[in[i]], b = T[in[i + 1]], c = T[in[i + 2]], d = T[in[i + 3]];
    if ((a | b | c | d) >= 0) {
      out[o++] = (uint8_t)((a << 2) | (b >> 4));
      out[o++] = (uint8_t)((b << 4) | (c >> 2));
      out[o++] = (uint8_t)((c << 6) | d);
    } else {
      if (a < 0 || b < 0) return -1;
      if (in[i + 2] == '=' && in[i + 3] == '=') {
        if (b & 15) return -1;
        out[o++] = (uint8_t)((a << 2) | (b >> 4));
      } else if (in[i + 3] == '=' && c >= 0) {
        if (c & 3) return -1;
        out[o++] = (uint8_t)((a << 2) | (b >> 4));
        out[o++] = (uint8_t)((b << 4) | (c >> 2));
      } else return -1;
    }
  }
  return (long)o;
}

}  // namespace zkb
