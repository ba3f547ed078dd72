// regexc.hpp — host regex compiler: pattern (Rust `regex` syntax subset, Unicode + UTF-8 mode as
// regex_automata::dfa::regex::Regex::new uses) -> forward and reverse dense DFAs in the ZDF1 layout
// (include/zkemail_b200.h).  Stands in for helpers/src/regex.rs:7-14 (`create_dfa`), which needs
// the regex-automata crate; the tables it emits are what `DFA.fwd` / `DFA.bwd`
// (core/src/structs.rs:16-19) carry in this engine and what the DFA kernel scans.
//
// Semantics reproduced: forward DFA = unanchored (lazy any-byte prefix) with
// leftmost-first match preference (lower-priority NFA threads are cut when a match state is
// reached), match states delayed by one byte, explicit end-of-input class; reverse DFA = anchored,
// built from the reversed expression with match-kind "all" (longest), so that a reverse scan from a
// match end yields the leftmost start.  Start states are selected by the look-behind byte class
// (NonWordByte, WordByte, Text, LineLF, LineCR, Custom) exactly as regex-automata's start table.
//
// Supported syntax: literals, escapes (\n \r \t \f \v \a \xHH \x{H..} \u{H..} \uHHHH and escaped
// punctuation), `.`, classes with ranges / negation / POSIX names / nested \d \w \s, \d \w \s \D \W \S
// (Unicode definitions in (?u) mode from generated tables, ASCII under (?-u)), groups (capturing, non-capturing,
// named), alternation, greedy and lazy * + ? {m} {m,} {m,n}, anchors ^ $ \A \z, flags i m s U u ((?i) = Unicode
// simple case folding, ASCII-only under (?-u)).
// ASCII word boundaries ((?-u:\b), (?-u:\B)) are resolved by the determiniser (one "previous byte was a word byte"
// bit per state, as regex-automata does).
// \p{..} / \P{..} with Unicode general categories (Lu, L, Letter, gc=Nd, ...), Any and ASCII.
// Rejected (ZKB_E_REGEX): Unicode word boundaries (no dense DFA exists for them), Unicode scripts and other
// non-category properties, class set operations, back-references, the x flag.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "unicode_tables.hpp"

namespace zkb {
namespace rx {

typedef std::pair<uint32_t, uint32_t> Range;  // inclusive
typedef std::vector<Range> RangeSet;

enum Look { LOOK_START_TEXT = 1, LOOK_END_TEXT = 2, LOOK_START_LINE = 4, LOOK_END_LINE = 8,
            LOOK_WORD_ASCII = 16, LOOK_NOT_WORD_ASCII = 32 };   // (?-u:\b) (?-u:\B)
inline bool is_word_byte(int b) { return (b >= '0' && b <= '9') || (b >= 'A' && b <= 'Z') || (b >= 'a' && b <= 'z') || b == '_'; }

inline void normalize(RangeSet& r) {
  std::sort(r.begin(), r.end());
  RangeSet o;
  for (auto& x : r) {
    if (!o.empty() && x.first <= o.back().second + 1 && o.back().second != 0xFFFFFFFFu) {
      o.back().second = std::max(o.back().second, x.second);
    } else o.push_back(x);
  }
  r.swap(o);
}
inline RangeSet negate(const RangeSet& in, uint32_t maxv) {
  RangeSet r = in, o;
  normalize(r);
  uint32_t next = 0;
  bool open = true;
  for (auto& x : r) {
    if (x.first > next) o.push_back(Range(next, x.first - 1));
    if (x.second >= maxv) { open = false; break; }
    next = x.second + 1;
  }
  if (open && next <= maxv) o.push_back(Range(next, maxv));
  return o;
}

// ---------------------------------------------------------------- HIR
struct Node {
  enum Kind { EMPTY, CLASS, CONCAT, ALT, REPEAT, LOOK } kind = EMPTY;
  RangeSet ranges;        // CLASS: code points (unicode=true) or bytes
  bool unicode = true;    // CLASS: ranges are Unicode scalar values encoded as UTF-8
  std::vector<int> kids;  // CONCAT / ALT / REPEAT (one kid)
  uint32_t min = 0, max = 0;  // REPEAT; max == INF for unbounded
  bool greedy = true;
  int look = 0;
};
static const uint32_t INF = 0xFFFFFFFFu;

struct Flags {
  bool i = false, m = false, s = false, U = false, u = true;
};

struct Parser {
  const uint8_t* p;
  size_t n, pos = 0;
  std::vector<Node> nodes;
  std::string err;
  int depth = 0;
  bool cur_unicode = true;   // the u flag in effect where an escape is being parsed
  bool cur_icase = false;    // the i flag in effect there

  Parser(const char* s, size_t len) : p((const uint8_t*)s), n(len) {}
  int add(const Node& nd) { nodes.push_back(nd); return (int)nodes.size() - 1; }
  bool fail(const std::string& m) { if (err.empty()) err = m + " at offset " + std::to_string(pos); return false; }
  bool eof() const { return pos >= n; }

  // decode one UTF-8 scalar of the pattern text
  bool next_char(uint32_t& c) {
    if (pos >= n) return fail("unexpected end of pattern");
    uint8_t b = p[pos];
    int w = b < 0x80 ? 1 : (b >> 5) == 6 ? 2 : (b >> 4) == 14 ? 3 : (b >> 3) == 30 ? 4 : 0;
    if (!w || pos + w > n) return fail("pattern is not valid UTF-8");
    c = w == 1 ? b : b & (0xFF >> (w + 1));
    for (int k = 1; k < w; k++) {
      if ((p[pos + k] & 0xC0) != 0x80) return fail("pattern is not valid UTF-8");
      c = (c << 6) | (p[pos + k] & 0x3F);
    }
    pos += w;
    return true;
  }
  // Simple case folding (Unicode CaseFolding C + S, tables from tools/gen_unicode_tables.py): every code point of
  // the set drags in the other members of its folding class (k K U+212A; s S U+017F; sigma, final sigma, Sigma; ...).
  static const std::vector<std::vector<uint32_t>>& fold_classes() {
    static const std::vector<std::vector<uint32_t>> classes = [] {
      std::map<uint32_t, std::vector<uint32_t>> by_fold;
      for (size_t i = 0; i < UNI_FOLD_N; i++) by_fold[UNI_FOLD[i][1]].push_back(UNI_FOLD[i][0]);
      std::vector<std::vector<uint32_t>> out;
      for (auto& kv : by_fold) {
        std::vector<uint32_t> c = kv.second;
        c.push_back(kv.first);
        std::sort(c.begin(), c.end());
        out.push_back(c);
      }
      return out;
    }();
    return classes;
  }
  static void fold_case(RangeSet& r, bool unicode) {
    normalize(r);
    if (!unicode) {   // (?i-u): bytes, ASCII letters only
      RangeSet add;
      for (auto& x : r) {
        uint32_t lo = std::max<uint32_t>(x.first, 'a'), hi = std::min<uint32_t>(x.second, 'z');
        if (lo <= hi) add.push_back(Range(lo - 32, hi - 32));
        lo = std::max<uint32_t>(x.first, 'A'); hi = std::min<uint32_t>(x.second, 'Z');
        if (lo <= hi) add.push_back(Range(lo + 32, hi + 32));
      }
      r.insert(r.end(), add.begin(), add.end());
      normalize(r);
      return;
    }
    auto has = [&](uint32_t c) {
      size_t lo = 0, hi = r.size();
      while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (r[mid].second < c) lo = mid + 1; else hi = mid;
      }
      return lo < r.size() && r[lo].first <= c;
    };
    RangeSet add;
    for (auto& cls : fold_classes()) {
      bool any = false;
      for (uint32_t c : cls) if (has(c)) { any = true; break; }
      if (any) for (uint32_t c : cls) add.push_back(Range(c, c));
    }
    r.insert(r.end(), add.begin(), add.end());
    normalize(r);
  }
  int make_class(RangeSet r, const Flags& f, bool negated) {
    if (f.i) {
      fold_case(r, f.u);
      if (!f.u) {  // byte mode: nothing above 0xFF
        RangeSet t;
        for (auto& x : r) if (x.first <= 0xFF) t.push_back(Range(x.first, std::min<uint32_t>(x.second, 0xFF)));
        r.swap(t);
      }
    }
    normalize(r);
    if (negated) r = negate(r, f.u ? 0x10FFFF : 0xFF);
    if (f.u) {  // drop surrogates
      RangeSet t;
      for (auto& x : r) {
        if (x.second < 0xD800 || x.first > 0xDFFF) { t.push_back(x); continue; }
        if (x.first < 0xD800) t.push_back(Range(x.first, 0xD7FF));
        if (x.second > 0xDFFF) t.push_back(Range(0xE000, x.second));
      }
      r.swap(t);
    } else {
      for (auto& x : r) if (x.second > 0x7F) { fail("pattern can match invalid UTF-8 (byte class above 0x7F under (?-u))"); return -1; }
    }
    Node nd;
    nd.kind = Node::CLASS; nd.ranges = r; nd.unicode = f.u;
    return add(nd);
  }
  // \d \w \s: Unicode definitions in (?u) mode (regex-syntax: \p{Nd}, \p{Alphabetic}+M+Nd+Pc+Join_Control,
  // \p{White_Space}; tables generated from Unicode 15.0, see tools/gen_unicode_tables.py), ASCII under (?-u)
  static bool perl_class(uint32_t c, RangeSet& out, bool& neg, bool unicode) {
    neg = (c == 'D' || c == 'W' || c == 'S');
    if (unicode) {
      const uint32_t (*t)[2] = nullptr;
      size_t n = 0;
      switch (c | 32) {
        case 'd': t = UNI_DIGIT; n = UNI_DIGIT_N; break;
        case 'w': t = UNI_WORD; n = UNI_WORD_N; break;
        case 's': t = UNI_SPACE; n = UNI_SPACE_N; break;
        default: return false;
      }
      out.clear();
      for (size_t i = 0; i < n; i++) out.push_back(Range(t[i][0], t[i][1]));
      return true;
    }
    switch (c | 32) {
      case 'd': out = {Range('0', '9')}; return true;
      case 'w': out = {Range('0', '9'), Range('A', 'Z'), Range('_', '_'), Range('a', 'z')}; return true;
      case 's': out = {Range('\t', '\r'), Range(' ', ' ')}; return true;
    }
    return false;
  }
  // \p{..} / \P{..} / \pL: Unicode general categories (two-letter values, their one-letter groups, long names,
  // gc= / General_Category= prefixes, \p{^..} negation, Any, ASCII).  Scripts and other properties are rejected:
  // the tables come from Python's unicodedata, which only exposes the general category.
  bool unicode_property(bool negated, RangeSet& out, bool& neg) {
    if (!cur_unicode) return fail("Unicode class not allowed under (?-u)");
    if (cur_icase) return fail("Unicode class under (?i) is not supported (no case-folding tables for it)");
    std::string name;
    if (pos < n && p[pos] == '{') {
      size_t a = ++pos;
      while (pos < n && p[pos] != '}') pos++;
      if (pos >= n) return fail("unclosed Unicode class");
      name.assign((const char*)p + a, pos - a);
      pos++;
    } else if (pos < n) {
      name.assign(1, (char)p[pos++]);
    } else return fail("incomplete Unicode class");
    bool inner_neg = false;
    std::string k;
    for (char c : name) {
      if (c == ' ' || c == '_' || c == '-') continue;
      k.push_back((char)((c >= 'A' && c <= 'Z') ? c + 32 : c));
    }
    if (!k.empty() && k[0] == '^') { inner_neg = true; k.erase(0, 1); }
    for (const char* pre : {"generalcategory=", "generalcategory:", "gc=", "gc:"})
      if (k.compare(0, strlen(pre), pre) == 0) { k.erase(0, strlen(pre)); break; }
    static const struct { const char* longname; const char* cats; } NAMES[] = {
        {"letter", "lu ll lt lm lo"}, {"l", "lu ll lt lm lo"}, {"casedletter", "lu ll lt"}, {"lc", "lu ll lt"},
        {"uppercaseletter", "lu"}, {"lowercaseletter", "ll"}, {"titlecaseletter", "lt"}, {"modifierletter", "lm"}, {"otherletter", "lo"},
        {"mark", "mn mc me"}, {"m", "mn mc me"}, {"combiningmark", "mn mc me"}, {"nonspacingmark", "mn"}, {"spacingmark", "mc"}, {"enclosingmark", "me"},
        {"number", "nd nl no"}, {"n", "nd nl no"}, {"decimalnumber", "nd"}, {"digit", "nd"}, {"letternumber", "nl"}, {"othernumber", "no"},
        {"punctuation", "pc pd ps pe pi pf po"}, {"p", "pc pd ps pe pi pf po"}, {"punct", "pc pd ps pe pi pf po"},
        {"connectorpunctuation", "pc"}, {"dashpunctuation", "pd"}, {"openpunctuation", "ps"}, {"closepunctuation", "pe"},
        {"initialpunctuation", "pi"}, {"finalpunctuation", "pf"}, {"otherpunctuation", "po"},
        {"symbol", "sm sc sk so"}, {"s", "sm sc sk so"}, {"mathsymbol", "sm"}, {"currencysymbol", "sc"}, {"modifiersymbol", "sk"}, {"othersymbol", "so"},
        {"separator", "zs zl zp"}, {"z", "zs zl zp"}, {"spaceseparator", "zs"}, {"lineseparator", "zl"}, {"paragraphseparator", "zp"},
        {"other", "cc cf co cn"}, {"c", "cc cf co cn"}, {"control", "cc"}, {"cntrl", "cc"}, {"format", "cf"}, {"privateuse", "co"}, {"unassigned", "cn"},
    };
    std::string cats;
    if (k == "any") { out = {Range(0, 0x10FFFF)}; neg = negated != inner_neg; return true; }
    if (k == "ascii") { out = {Range(0, 0x7F)}; neg = negated != inner_neg; return true; }
    for (auto& nm : NAMES) if (k == nm.longname) { cats = nm.cats; break; }
    if (cats.empty() && k.size() == 2) cats = k;
    out.clear();
    size_t a = 0;
    while (a < cats.size()) {
      size_t b = cats.find(' ', a);
      if (b == std::string::npos) b = cats.size();
      const std::string c2 = cats.substr(a, b - a);
      bool found = false;
      for (size_t i = 0; i < UNI_GC_N; i++)
        if (c2 == UNI_GC[i].name) {
          for (size_t j = 0; j < UNI_GC[i].n; j++) out.push_back(Range(UNI_GC[i].r[j][0], UNI_GC[i].r[j][1]));
          found = true;
          break;
        }
      if (!found) { out.clear(); break; }
      a = b + 1;
    }
    if (out.empty()) return fail("unsupported Unicode property (general categories only)");
    normalize(out);
    neg = negated != inner_neg;
    return true;
  }
  bool hex_digits(int count, uint32_t& v) {
    v = 0;
    for (int k = 0; k < count; k++) {
      if (pos >= n) return fail("incomplete hex escape");
      uint8_t c = p[pos++];
      int d = c >= '0' && c <= '9' ? c - '0' : (c | 32) >= 'a' && (c | 32) <= 'f' ? (c | 32) - 'a' + 10 : -1;
      if (d < 0) return fail("invalid hex digit");
      v = v * 16 + d;
    }
    return true;
  }
  bool hex_escape(uint32_t kind, uint32_t& v) {
    if (pos < n && p[pos] == '{') {
      pos++;
      v = 0;
      int cnt = 0;
      while (pos < n && p[pos] != '}') {
        uint8_t c = p[pos++];
        int d = c >= '0' && c <= '9' ? c - '0' : (c | 32) >= 'a' && (c | 32) <= 'f' ? (c | 32) - 'a' + 10 : -1;
        if (d < 0 || ++cnt > 8) return fail("invalid hex escape");
        v = v * 16 + d;
      }
      if (pos >= n || cnt == 0) return fail("invalid hex escape");
      pos++;
    } else if (!hex_digits(kind == 'x' ? 2 : kind == 'u' ? 4 : 8, v)) return false;
    if (v > 0x10FFFF || (v >= 0xD800 && v <= 0xDFFF)) return fail("hex escape is not a Unicode scalar value");
    return true;
  }
  // escape after '\'; returns: 0 literal in `c`, 1 class in `cls`/`neg`, 2 look in `look`
  int escape(uint32_t& c, RangeSet& cls, bool& neg, int& look, bool in_class) {
    uint32_t e;
    if (!next_char(e)) return -1;
    switch (e) {
      case 'n': c = '\n'; return 0;
      case 'r': c = '\r'; return 0;
      case 't': c = '\t'; return 0;
      case 'f': c = 0x0C; return 0;
      case 'v': c = 0x0B; return 0;
      case 'a': c = 0x07; return 0;
      case 'x': case 'u': case 'U': return hex_escape(e, c) ? 0 : -1;
      case 'd': case 'D': case 'w': case 'W': case 's': case 'S': perl_class(e, cls, neg, cur_unicode); return 1;
      case 'p': case 'P': return unicode_property(e == 'P', cls, neg) ? 1 : -1;
      case 'A': if (in_class) break; look = LOOK_START_TEXT; return 2;
      case 'z': if (in_class) break; look = LOOK_END_TEXT; return 2;
      case 'b': case 'B':
        if (in_class) break;
        // DFARegex::new cannot build a dense DFA with Unicode word boundaries either (helpers/src/regex.rs:20 would
        // return Err); the ASCII ones are look-arounds the determiniser resolves with one bit of state
        if (cur_unicode) { fail("Unicode word boundary cannot be compiled to a DFA; write (?-u:\\b)"); return -1; }
        look = e == 'b' ? LOOK_WORD_ASCII : LOOK_NOT_WORD_ASCII;
        return 2;
      default: break;
    }
    if (e < 0x80 && !((e >= '0' && e <= '9') || ((e | 32) >= 'a' && (e | 32) <= 'z')) && e != '<' && e != '>') {
      c = e;  // escaped punctuation
      return 0;
    }
    fail("unsupported escape sequence");
    return -1;
  }
  bool posix_class(RangeSet& out, bool& neg) {  // at "[:"
    size_t save = pos;
    pos += 2;
    neg = false;
    if (pos < n && p[pos] == '^') { neg = true; pos++; }
    size_t a = pos;
    while (pos < n && p[pos] != ':') pos++;
    if (pos + 1 >= n || p[pos + 1] != ']') { pos = save; return false; }
    std::string nm((const char*)p + a, pos - a);
    pos += 2;
    static const struct { const char* n; RangeSet r; } T[] = {
        {"alnum", {Range('0', '9'), Range('A', 'Z'), Range('a', 'z')}},
        {"alpha", {Range('A', 'Z'), Range('a', 'z')}},
        {"ascii", {Range(0, 0x7F)}},
        {"blank", {Range('\t', '\t'), Range(' ', ' ')}},
        {"cntrl", {Range(0, 0x1F), Range(0x7F, 0x7F)}},
        {"digit", {Range('0', '9')}},
        {"graph", {Range('!', '~')}},
        {"lower", {Range('a', 'z')}},
        {"print", {Range(' ', '~')}},
        {"punct", {Range('!', '/'), Range(':', '@'), Range('[', '`'), Range('{', '~')}},
        {"space", {Range('\t', '\r'), Range(' ', ' ')}},
        {"upper", {Range('A', 'Z')}},
        {"word", {Range('0', '9'), Range('A', 'Z'), Range('_', '_'), Range('a', 'z')}},
        {"xdigit", {Range('0', '9'), Range('A', 'F'), Range('a', 'f')}},
    };
    for (auto& t : T) if (nm == t.n) { out = t.r; return true; }
    pos = save;
    return false;
  }
  int parse_class(const Flags& f) {  // at '['
    cur_unicode = f.u;
    cur_icase = f.i;
    pos++;
    bool negated = false;
    if (pos < n && p[pos] == '^') { negated = true; pos++; }
    RangeSet r;
    bool first = true;
    for (;;) {
      if (pos >= n) { fail("unclosed character class"); return -1; }
      if (p[pos] == ']' && !first) { pos++; break; }
      first = false;
      if (p[pos] == '[') {
        RangeSet pr;
        bool pneg;
        if (pos + 1 < n && p[pos + 1] == ':' && posix_class(pr, pneg)) {
          if (pneg) pr = negate(pr, f.u ? 0x10FFFF : 0xFF);
          r.insert(r.end(), pr.begin(), pr.end());
          continue;
        }
        fail("nested character classes are not supported");
        return -1;
      }
      if ((p[pos] == '&' && pos + 1 < n && p[pos + 1] == '&') || (p[pos] == '~' && pos + 1 < n && p[pos + 1] == '~') ||
          (p[pos] == '-' && pos + 1 < n && p[pos + 1] == '-')) { fail("class set operations are not supported"); return -1; }
      uint32_t lo;
      if (p[pos] == '\\') {
        pos++;
        RangeSet cls; bool neg; int look;
        int k = escape(lo, cls, neg, look, true);
        if (k < 0) return -1;
        if (k == 1) {
          if (neg) cls = negate(cls, f.u ? 0x10FFFF : 0xFF);
          r.insert(r.end(), cls.begin(), cls.end());
          continue;
        }
      } else if (!next_char(lo)) return -1;
      uint32_t hi = lo;
      if (pos + 1 < n && p[pos] == '-' && p[pos + 1] != ']') {
        pos++;
        if (p[pos] == '\\') {
          pos++;
          RangeSet cls; bool neg; int look;
          int k = escape(hi, cls, neg, look, true);
          if (k != 0) { fail("invalid class range"); return -1; }
        } else if (!next_char(hi)) return -1;
        if (hi < lo) { fail("invalid class range (start > end)"); return -1; }
      }
      r.push_back(Range(lo, hi));
    }
    return make_class(r, f, negated);
  }
  bool parse_uint(uint32_t& v) {
    size_t a = pos;
    uint64_t x = 0;
    while (pos < n && p[pos] >= '0' && p[pos] <= '9') { x = x * 10 + (p[pos] - '0'); if (x > 1000) return fail("repetition count too large"); pos++; }
    if (pos == a) return false;
    v = (uint32_t)x;
    return true;
  }
  // atom with optional repetition suffixes
  int parse_repeat(Flags& f) {
    int atom = parse_atom(f);
    if (atom < 0) return -1;
    for (;;) {
      if (pos >= n) break;
      uint32_t mn = 0, mx = 0;
      uint8_t c = p[pos];
      if (c == '*') { mn = 0; mx = INF; pos++; }
      else if (c == '+') { mn = 1; mx = INF; pos++; }
      else if (c == '?') { mn = 0; mx = 1; pos++; }
      else if (c == '{') {
        size_t save = pos;
        pos++;
        if (!parse_uint(mn)) { if (!err.empty()) return -1; pos = save; fail("invalid repetition"); return -1; }
        mx = mn;
        if (pos < n && p[pos] == ',') {
          pos++;
          if (pos < n && p[pos] == '}') mx = INF;
          else if (!parse_uint(mx)) { fail("invalid repetition"); return -1; }
        }
        if (pos >= n || p[pos] != '}') { fail("invalid repetition"); return -1; }
        pos++;
        if (mx != INF && mx < mn) { fail("invalid repetition range"); return -1; }
      } else break;
      // (a repetition of an assertion - `^*`, `(\b)+` - is legal in regex-syntax: only an empty expression or a flag
      // group in front of the operator is "repetition operator missing expression")
      bool greedy = true;
      if (pos < n && p[pos] == '?') { greedy = false; pos++; }
      if (f.U) greedy = !greedy;
      Node nd;
      nd.kind = Node::REPEAT; nd.kids = {atom}; nd.min = mn; nd.max = mx; nd.greedy = greedy;
      atom = add(nd);
    }
    return atom;
  }
  bool parse_flags(Flags& f, bool& group_follows) {  // after "(?"; stops after ':' or ')'
    bool neg = false, any = false;
    for (;;) {
      if (pos >= n) return fail("unclosed group");
      uint8_t c = p[pos++];
      if (c == ':') { group_follows = true; return true; }
      if (c == ')') { group_follows = false; return any ? true : fail("empty flags"); }
      if (c == '-') { if (neg) return fail("invalid flags"); neg = true; continue; }
      any = true;
      switch (c) {
        case 'i': f.i = !neg; break;
        case 'm': f.m = !neg; break;
        case 's': f.s = !neg; break;
        case 'U': f.U = !neg; break;
        case 'u': f.u = !neg; break;
        default: return fail("unsupported flag");
      }
    }
  }
  int parse_atom(Flags& f) {
    uint8_t c = p[pos];
    if (c == '(') {
      pos++;
      Flags inner = f;
      if (pos < n && p[pos] == '?') {
        if (pos + 1 < n && (p[pos + 1] == 'P' || p[pos + 1] == '<')) {  // named group
          pos += p[pos + 1] == 'P' ? 2 : 1;
          if (pos >= n || p[pos] != '<') { fail("invalid group name"); return -1; }
          pos++;
          size_t a = pos;
          while (pos < n && p[pos] != '>') pos++;
          if (pos >= n || pos == a) { fail("invalid group name"); return -1; }
          pos++;
        } else {
          pos++;
          bool group_follows;
          if (!parse_flags(inner, group_follows)) return -1;
          if (!group_follows) {  // (?flags) applies to the rest of the enclosing group
            f = inner;
            Node e;
            return add(e);
          }
        }
      }
      if (++depth > 200) { fail("nesting too deep"); return -1; }
      int sub = parse_alt(inner);
      depth--;
      if (sub < 0) return -1;
      if (pos >= n || p[pos] != ')') { fail("unclosed group"); return -1; }
      pos++;
      return sub;
    }
    if (c == '[') return parse_class(f);
    if (c == '.') {
      pos++;
      RangeSet r;
      if (!f.s) r.push_back(Range('\n', '\n'));
      Flags g = f;
      g.i = false;
      return make_class(r, g, true);
    }
    if (c == '^' || c == '$') {
      pos++;
      Node nd;
      nd.kind = Node::LOOK;
      nd.look = c == '^' ? (f.m ? LOOK_START_LINE : LOOK_START_TEXT) : (f.m ? LOOK_END_LINE : LOOK_END_TEXT);
      return add(nd);
    }
    if (c == '\\') {
      pos++;
      cur_unicode = f.u;
      cur_icase = f.i;
      uint32_t lit; RangeSet cls; bool neg; int look;
      int k = escape(lit, cls, neg, look, false);
      if (k < 0) return -1;
      if (k == 1) { Flags g = f; g.i = false; return make_class(cls, g, neg); }
      if (k == 2) { Node nd; nd.kind = Node::LOOK; nd.look = look; return add(nd); }
      return make_class({Range(lit, lit)}, f, false);
    }
    if (c == '*' || c == '+' || c == '?') { fail("repetition operator missing expression"); return -1; }
    if (c == '{') { fail("repetition operator missing expression"); return -1; }
    uint32_t lit;
    if (!next_char(lit)) return -1;
    return make_class({Range(lit, lit)}, f, false);
  }
  int parse_concat(Flags& f) {
    Node nd;
    nd.kind = Node::CONCAT;
    while (pos < n && p[pos] != '|' && p[pos] != ')') {
      int k = parse_repeat(f);
      if (k < 0) return -1;
      nd.kids.push_back(k);
    }
    if (nd.kids.size() == 1) return nd.kids[0];
    if (nd.kids.empty()) nd.kind = Node::EMPTY;
    return add(nd);
  }
  int parse_alt(Flags f) {
    Node nd;
    nd.kind = Node::ALT;
    for (;;) {
      int k = parse_concat(f);
      if (k < 0) return -1;
      nd.kids.push_back(k);
      if (pos < n && p[pos] == '|') { pos++; continue; }
      break;
    }
    if (nd.kids.size() == 1) return nd.kids[0];
    return add(nd);
  }
  int parse() {
    Flags f;
    int root = parse_alt(f);
    if (root < 0) return -1;
    if (pos < n) { fail(p[pos] == ')' ? "unopened group" : "unexpected character"); return -1; }
    return root;
  }
  bool can_be_empty(int id) const {
    const Node& nd = nodes[id];
    switch (nd.kind) {
      case Node::EMPTY: case Node::LOOK: return true;
      case Node::CLASS: return false;
      case Node::CONCAT: for (int k : nd.kids) if (!can_be_empty(k)) return false; return true;
      case Node::ALT: for (int k : nd.kids) if (can_be_empty(k)) return true; return false;
      case Node::REPEAT: return nd.min == 0 || can_be_empty(nd.kids[0]);
    }
    return false;
  }
};

// ---------------------------------------------------------------- UTF-8 range sequences
struct ByteSeq { int len; uint8_t lo[4], hi[4]; };
inline int utf8_encode(uint32_t c, uint8_t* o) {
  if (c < 0x80) { o[0] = (uint8_t)c; return 1; }
  if (c < 0x800) { o[0] = 0xC0 | (c >> 6); o[1] = 0x80 | (c & 0x3F); return 2; }
  if (c < 0x10000) { o[0] = 0xE0 | (c >> 12); o[1] = 0x80 | ((c >> 6) & 0x3F); o[2] = 0x80 | (c & 0x3F); return 3; }
  o[0] = 0xF0 | (c >> 18); o[1] = 0x80 | ((c >> 12) & 0x3F); o[2] = 0x80 | ((c >> 6) & 0x3F); o[3] = 0x80 | (c & 0x3F);
  return 4;
}
inline void utf8_sequences(uint32_t lo, uint32_t hi, std::vector<ByteSeq>& out) {
  std::vector<Range> st;
  st.push_back(Range(lo, hi));
  while (!st.empty()) {
    uint32_t s = st.back().first, e = st.back().second;
    st.pop_back();
    for (;;) {
      if (s < 0xE000 && e > 0xD7FF) {  // split around the surrogate gap
        if (e >= 0xE000) st.push_back(Range(0xE000, e));
        if (s > 0xD7FF) { s = 1; e = 0; break; }
        e = 0xD7FF;
      }
      if (e < s) break;
      bool again = false;
      static const uint32_t MAXV[3] = {0x7F, 0x7FF, 0xFFFF};
      for (int i = 0; i < 3; i++)
        if (s <= MAXV[i] && MAXV[i] < e) { st.push_back(Range(MAXV[i] + 1, e)); e = MAXV[i]; again = true; break; }
      if (again) continue;
      if (e <= 0x7F) break;
      for (int i = 1; i < 4; i++) {
        uint32_t m = (1u << (6 * i)) - 1;
        if ((s & ~m) != (e & ~m)) {
          if ((s & m) != 0) { st.push_back(Range((s | m) + 1, e)); e = s | m; again = true; break; }
          if ((e & m) != m) { st.push_back(Range(e & ~m, e)); e = (e & ~m) - 1; again = true; break; }
        }
      }
      if (again) continue;
      break;
    }
    if (e < s) continue;
    ByteSeq q;
    uint8_t a[4], b[4];
    int la = utf8_encode(s, a), lb = utf8_encode(e, b);
    (void)lb;
    q.len = la;
    for (int i = 0; i < la; i++) { q.lo[i] = a[i]; q.hi[i] = b[i]; }
    out.push_back(q);
  }
}

// ---------------------------------------------------------------- Thompson NFA
struct NState {
  enum T { RANGE, UNION, LOOKS, MATCH } t = UNION;
  uint8_t lo = 0, hi = 0;
  int look = 0;
  int next = -1;
  std::vector<int> alts;  // UNION (in priority order)
};
struct Nfa {
  std::vector<NState> st;
  int start_anchored = -1, start_unanchored = -1;
  int looks_any = 0;
  bool overflow = false;
  int add(const NState& s) {
    if (st.size() > 200000) { overflow = true; return 0; }
    st.push_back(s);
    return (int)st.size() - 1;
  }
  int range(uint8_t lo, uint8_t hi, int next) { NState s; s.t = NState::RANGE; s.lo = lo; s.hi = hi; s.next = next; return add(s); }
  int uni(const std::vector<int>& a) { NState s; s.t = NState::UNION; s.alts = a; return add(s); }
};

struct NfaBuilder {
  const std::vector<Node>& nodes;
  Nfa& nfa;
  bool reverse;
  NfaBuilder(const std::vector<Node>& nd, Nfa& n, bool rev) : nodes(nd), nfa(n), reverse(rev) {}

  // shortest match length in bytes, saturated (regex-syntax: Properties::minimum_len; only "is it zero" matters here)
  uint32_t min_len(int id) const {
    const Node& nd = nodes[id];
    switch (nd.kind) {
      case Node::EMPTY: case Node::LOOK: return 0;
      case Node::CLASS: return 1;
      case Node::CONCAT: { uint64_t a = 0; for (int k : nd.kids) a += min_len(k); return a > 0x7fffffffu ? 0x7fffffffu : (uint32_t)a; }
      case Node::ALT: { uint32_t m = 0x7fffffffu; for (int k : nd.kids) m = std::min(m, min_len(k)); return nd.kids.empty() ? 0 : m; }
      case Node::REPEAT: { uint64_t a = (uint64_t)nd.min * min_len(nd.kids[0]); return a > 0x7fffffffu ? 0x7fffffffu : (uint32_t)a; }
    }
    return 0;
  }

  int compile(int id, int next) {
    if (nfa.overflow) return next;
    const Node& nd = nodes[id];
    switch (nd.kind) {
      case Node::EMPTY: return next;
      case Node::LOOK: {
        int lk = nd.look;
        if (reverse && !(lk & (LOOK_WORD_ASCII | LOOK_NOT_WORD_ASCII)))   // word boundaries read the same in both directions
          lk = lk == LOOK_START_TEXT ? LOOK_END_TEXT : lk == LOOK_END_TEXT ? LOOK_START_TEXT : lk == LOOK_START_LINE ? LOOK_END_LINE : LOOK_START_LINE;
        NState s; s.t = NState::LOOKS; s.look = lk; s.next = next;
        nfa.looks_any |= lk;
        return nfa.add(s);
      }
      case Node::CLASS: {
        std::vector<int> alts;
        if (!nd.unicode) {
          for (auto& r : nd.ranges) alts.push_back(nfa.range((uint8_t)r.first, (uint8_t)r.second, next));
        } else {
          std::vector<ByteSeq> seqs;
          for (auto& r : nd.ranges) utf8_sequences(r.first, r.second, seqs);
          for (auto& q : seqs) {
            int nx = next;
            if (!reverse) for (int i = q.len - 1; i >= 0; i--) nx = nfa.range(q.lo[i], q.hi[i], nx);
            else for (int i = 0; i < q.len; i++) nx = nfa.range(q.lo[i], q.hi[i], nx);
            alts.push_back(nx);
          }
        }
        if (alts.empty()) { NState s; s.t = NState::UNION; return nfa.add(s); }  // matches nothing
        if (alts.size() == 1) return alts[0];
        return nfa.uni(alts);
      }
      case Node::CONCAT: {
        int nx = next;
        if (!reverse) for (int i = (int)nd.kids.size() - 1; i >= 0; i--) nx = compile(nd.kids[i], nx);
        else for (size_t i = 0; i < nd.kids.size(); i++) nx = compile(nd.kids[i], nx);
        return nx;
      }
      case Node::ALT: {
        std::vector<int> alts;
        for (int k : nd.kids) alts.push_back(compile(k, next));
        return nfa.uni(alts);
      }
      case Node::REPEAT: {
        // The shapes of regex-automata's Thompson compiler (c_at_least / c_bounded / c_exactly), not just the language:
        // leftmost-first preference is decided by the order in which the epsilon closure meets the states, and that
        // depends on how many copies of the sub-expression exist and where the back edge goes.
        //   e{n,}, n >= 1 : e^(n-1) then ONE copy of e with a union behind it (back to that copy | exit)
        //   e*, e cannot match the empty string : union(e -> union | exit)
        //   e*, e can match the empty string    : (e+)?  - with the plain loop the empty path through e would come back
        //        to the loop head, be dropped as already visited, and e's other alternatives would outrank the exit,
        //        so that (|a)* on "aa" took "aa" where the crate (and Perl) take "" (rust-lang/regex issue 779)
        //   e{m,n} : e^m then nested optionals e(e(e)?)?
        int sub = nd.kids[0];
        int t = next;
        uint32_t copies = nd.min;
        if (nd.max == INF) {
          int u = nfa.uni({});
          int body = compile(sub, u);
          if (nd.greedy) nfa.st[u].alts = {body, next}; else nfa.st[u].alts = {next, body};
          if (nd.min == 0) {
            if (min_len(sub) > 0) t = u;
            else t = nd.greedy ? nfa.uni({body, next}) : nfa.uni({next, body});
          } else {
            t = body;          // the copy in front of the union is the last of the n mandatory ones
            copies = nd.min - 1;
          }
        } else {
          for (uint32_t i = nd.min; i < nd.max; i++) {
            int body = compile(sub, t);
            t = nd.greedy ? nfa.uni({body, next}) : nfa.uni({next, body});
          }
        }
        for (uint32_t i = 0; i < copies; i++) t = compile(sub, t);
        return t;
      }
    }
    return next;
  }
};

// ---------------------------------------------------------------- determinisation
struct DState {
  std::vector<int> ids;  // NFA states (RANGE / LOOKS / MATCH) in priority order
  int look_have = 0, look_need = 0;
  bool is_match = false;
  bool from_word = false;   // the byte that led here is an ASCII word byte (tracked only for patterns with \b / \B)
};
struct Dfa {
  uint32_t n_classes = 0;  // including EOI
  uint8_t class_map[256];
  std::vector<std::vector<uint32_t>> trans;  // [state][class]
  std::vector<bool> is_match;
  uint32_t start[12];
};

struct Determinizer {
  const Nfa& nfa;
  bool leftmost_first;  // false: match-kind all
  std::vector<DState> states;
  std::map<std::vector<int>, uint32_t> index;
  std::vector<int> mark;  // epoch per NFA state
  int epoch = 0;
  Determinizer(const Nfa& n, bool lf) : nfa(n), leftmost_first(lf), mark(n.st.size(), 0) {}

  void closure(int from, int look_have, std::vector<int>& out) {
    std::vector<int> stack;
    stack.push_back(from);
    while (!stack.empty()) {
      int id = stack.back();
      stack.pop_back();
      for (;;) {
        if (mark[id] == epoch) break;
        mark[id] = epoch;
        const NState& s = nfa.st[id];
        if (s.t == NState::RANGE || s.t == NState::MATCH) { out.push_back(id); break; }
        if (s.t == NState::LOOKS) {
          out.push_back(id);
          if ((look_have & s.look) == 0) break;
          id = s.next;
          continue;
        }
        if (s.alts.empty()) break;
        for (size_t k = s.alts.size(); k-- > 1;) stack.push_back(s.alts[k]);
        id = s.alts[0];
      }
    }
  }
  uint32_t intern(DState& d) {
    d.look_need = 0;
    for (int id : d.ids) if (nfa.st[id].t == NState::LOOKS) d.look_need |= nfa.st[id].look;
    if (!d.look_need) d.look_have = 0;
    std::vector<int> key;
    key.reserve(d.ids.size() + 2);
    if (!(nfa.looks_any & (LOOK_WORD_ASCII | LOOK_NOT_WORD_ASCII))) d.from_word = false;
    key.push_back((d.is_match ? 1 : 0) | (d.from_word ? 2 : 0));
    key.push_back(d.look_have);
    key.insert(key.end(), d.ids.begin(), d.ids.end());
    auto it = index.find(key);
    if (it != index.end()) return it->second;
    uint32_t id = (uint32_t)states.size();
    states.push_back(d);
    index.emplace(std::move(key), id);
    return id;
  }
  uint32_t start_state(int nfa_start, int look_have, bool from_word) {
    DState d;
    d.look_have = look_have;
    d.from_word = from_word;
    epoch++;
    closure(nfa_start, look_have, d.ids);
    return intern(d);
  }
  // unit: 0..255 byte, 256 = end of input
  // src must be a copy of states[from] (intern() may grow `states`)
  uint32_t next(const DState& src, int unit) {
    std::vector<int> cur = src.ids;
    if (src.look_need) {
      int have = src.look_have;
      if (unit == 256) have |= LOOK_END_TEXT | LOOK_END_LINE;
      if (unit == '\n') have |= LOOK_END_LINE;
      if (nfa.looks_any & (LOOK_WORD_ASCII | LOOK_NOT_WORD_ASCII))   // between the byte behind and the byte ahead (EOI: not a word byte)
        have |= (src.from_word != (unit < 256 && is_word_byte(unit))) ? LOOK_WORD_ASCII : LOOK_NOT_WORD_ASCII;
      if ((have & ~src.look_have) & src.look_need) {
        std::vector<int> re;
        epoch++;
        for (int id : cur) closure(id, have, re);
        cur.swap(re);
      }
    }
    DState d;
    d.from_word = unit < 256 && is_word_byte(unit);
    if ((nfa.looks_any & (LOOK_START_LINE | LOOK_END_LINE)) && unit == '\n') d.look_have |= LOOK_START_LINE;
    epoch++;
    for (int id : cur) {
      const NState& s = nfa.st[id];
      if (s.t == NState::MATCH) {
        d.is_match = true;
        if (leftmost_first) break;
      } else if (s.t == NState::RANGE) {
        if (unit < 256 && unit >= s.lo && unit <= s.hi) closure(s.next, d.look_have, d.ids);
      }
    }
    return intern(d);
  }
};

inline bool determinize(const Nfa& nfa, bool leftmost_first, bool with_unanchored, Dfa& out, std::string& err) {
  // byte classes from range boundaries
  bool bound[257];
  memset(bound, 0, sizeof bound);
  bound[0] = true;
  for (auto& s : nfa.st)
    if (s.t == NState::RANGE) { bound[s.lo] = true; bound[(int)s.hi + 1] = true; }
  if (nfa.looks_any & (LOOK_START_LINE | LOOK_END_LINE)) { bound['\n'] = true; bound['\n' + 1] = true; }
  if (nfa.looks_any & (LOOK_WORD_ASCII | LOOK_NOT_WORD_ASCII))
    for (int b = 1; b <= 256; b++) if (is_word_byte(b - 1) != (b < 256 && is_word_byte(b))) bound[b] = true;
  std::vector<int> rep;
  int cls = -1;
  for (int b = 0; b < 256; b++) {
    if (bound[b]) { cls++; rep.push_back(b); }
    out.class_map[b] = (uint8_t)cls;
  }
  out.n_classes = (uint32_t)cls + 2;
  rep.push_back(256);

  Determinizer det(nfa, leftmost_first);
  { DState dead; det.intern(dead); }  // id 0
  // start kinds: NonWordByte, WordByte, Text, LineLF, LineCR, CustomLineTerminator
  static const int KIND_LOOK[6] = {0, 0, LOOK_START_TEXT | LOOK_START_LINE, LOOK_START_LINE, 0, 0};
  for (int k = 0; k < 6; k++) {
    uint32_t a = det.start_state(nfa.start_anchored, KIND_LOOK[k], k == 1);
    uint32_t u = with_unanchored ? det.start_state(nfa.start_unanchored, KIND_LOOK[k], k == 1) : a;
    out.start[k] = u;
    out.start[6 + k] = a;
  }
  for (uint32_t s = 0; s < det.states.size(); s++) {
    if (det.states.size() > 20000) { err = "DFA too large (more than 20000 states)"; return false; }
    std::vector<uint32_t> row(out.n_classes);
    const DState src = det.states[s];
    for (uint32_t c = 0; c < out.n_classes; c++) row[c] = det.next(src, rep[c]);
    out.trans.push_back(row);
  }
  out.is_match.resize(det.states.size());
  for (size_t s = 0; s < det.states.size(); s++) out.is_match[s] = det.states[s].is_match;
  return true;
}

// Moore partition refinement: merges states that no input can tell apart (same match flag now and after every
// byte-class / EOI sequence).  Search results are unchanged: the scan only observes transitions, match flags and
// the dead state.  States that can no longer reach a match collapse into the dead state, so scans also stop earlier.
// Smaller tables matter on the device: they stay in shared memory and qualify for 256-wide DIRECT rows.
inline void minimize(Dfa& d) {
  const size_t ns = d.trans.size(), nc = d.n_classes;
  if (ns <= 1) return;
  std::vector<uint32_t> cls(ns);
  for (size_t s = 0; s < ns; s++) cls[s] = d.is_match[s] ? 1u : 0u;
  size_t n_cls = 2;
  for (;;) {
    std::map<std::vector<uint32_t>, uint32_t> sig_id;
    std::vector<uint32_t> next_cls(ns);
    std::vector<uint32_t> sig(nc + 1);
    for (size_t s = 0; s < ns; s++) {
      sig[0] = cls[s];
      for (size_t c = 0; c < nc; c++) sig[c + 1] = cls[d.trans[s][c]];
      auto it = sig_id.find(sig);
      if (it == sig_id.end()) it = sig_id.emplace(sig, (uint32_t)sig_id.size()).first;
      next_cls[s] = it->second;
    }
    const size_t n_new = sig_id.size();
    cls.swap(next_cls);
    if (n_new == n_cls) break;
    n_cls = n_new;
  }
  // renumber so that the dead state's block is 0
  std::vector<uint32_t> remap(n_cls, 0xFFFFFFFFu);
  uint32_t next_id = 0;
  remap[cls[0]] = next_id++;
  for (size_t s = 0; s < ns; s++) if (remap[cls[s]] == 0xFFFFFFFFu) remap[cls[s]] = next_id++;
  std::vector<std::vector<uint32_t>> trans(n_cls, std::vector<uint32_t>(nc, 0));
  std::vector<bool> is_match(n_cls, false);
  std::vector<bool> done(n_cls, false);
  for (size_t s = 0; s < ns; s++) {
    const uint32_t b = remap[cls[s]];
    if (done[b]) continue;
    done[b] = true;
    is_match[b] = d.is_match[s];
    for (size_t c = 0; c < nc; c++) trans[b][c] = remap[cls[d.trans[s][c]]];
  }
  for (int i = 0; i < 12; i++) d.start[i] = remap[cls[d.start[i]]];
  d.trans.swap(trans);
  d.is_match.swap(is_match);
}

// ZDF1 serialisation: dead state 0, then match states as one contiguous id range, then the rest.
inline void emit_zdf(const Dfa& d, uint32_t flags, std::vector<uint8_t>& out) {
  const uint32_t ns = (uint32_t)d.trans.size();
  std::vector<uint32_t> remap(ns, 0);
  uint32_t next_id = 1, n_match = 0;
  for (uint32_t s = 1; s < ns; s++) if (d.is_match[s]) { remap[s] = next_id++; n_match++; }
  for (uint32_t s = 1; s < ns; s++) if (!d.is_match[s]) remap[s] = next_id++;
  out.assign(584 + (size_t)ns * d.n_classes * 4, 0);
  auto w32 = [&](size_t off, uint32_t v) { out[off] = (uint8_t)v; out[off + 1] = (uint8_t)(v >> 8); out[off + 2] = (uint8_t)(v >> 16); out[off + 3] = (uint8_t)(v >> 24); };
  w32(0, 0x3146445Au); w32(4, flags); w32(8, ns); w32(12, d.n_classes);
  if (n_match) { w32(16, 1); w32(20, n_match); } else { w32(16, 1); w32(20, 0); }
  for (int i = 0; i < 12; i++) w32(24 + 4 * i, remap[d.start[i]]);
  memcpy(out.data() + 72, d.class_map, 256);
  for (int b = 0; b < 256; b++) {
    uint8_t k = 0;  // NonWordByte
    if ((b >= '0' && b <= '9') || (b >= 'A' && b <= 'Z') || (b >= 'a' && b <= 'z') || b == '_') k = 1;
    else if (b == '\n') k = 3;
    else if (b == '\r') k = 4;
    out[328 + b] = k;
  }
  for (uint32_t s = 0; s < ns; s++)
    for (uint32_t c = 0; c < d.n_classes; c++) w32(584 + ((size_t)remap[s] * d.n_classes + c) * 4, remap[d.trans[s][c]]);
}

// pattern -> (fwd, bwd) ZDF1 tables.  false + err on rejection.
inline bool compile(const char* pattern, size_t len, std::vector<uint8_t>& fwd, std::vector<uint8_t>& bwd, std::string& err) {
  Parser ps(pattern, len);
  int root = ps.parse();
  if (root < 0) { err = ps.err.empty() ? "regex parse error" : ps.err; return false; }
  const bool has_empty = ps.can_be_empty(root);
  const uint32_t base_flags = 2u /* utf8 */ | (has_empty ? 4u : 0u);
  for (int rev = 0; rev < 2; rev++) {
    Nfa nfa;
    NState m; m.t = NState::MATCH;
    int match = nfa.add(m);
    NfaBuilder b(ps.nodes, nfa, rev != 0);
    nfa.start_anchored = b.compile(root, match);
    // unanchored prefix (?s-u:.)*? : prefer the pattern, else consume any byte and retry
    int loop = nfa.uni({});
    int any = nfa.range(0, 255, loop);
    nfa.st[loop].alts = {nfa.start_anchored, any};
    nfa.start_unanchored = loop;
    if (nfa.overflow) { err = "pattern too large (NFA state limit)"; return false; }
    Dfa d;
    if (!determinize(nfa, rev == 0, rev == 0, d, err)) return false;
    minimize(d);
    emit_zdf(d, base_flags | (rev ? 1u : 0u), rev ? bwd : fwd);
  }
  return true;
}

}  // namespace rx
}  // namespace zkb
