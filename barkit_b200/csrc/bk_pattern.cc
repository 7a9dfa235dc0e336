// Pattern front end: the host-side replacement of barkit-extract/src/pattern.rs.
//
//   find_fuzzy_runs      <- ADAPTER_PATTERN_REGEX            pattern.rs:11  (fancy-regex find_iter :97)
//   sequence_with_errors <- get_sequence_with_errors        pattern.rs:40-79
//   expand               <- get_pattern_with_errors         pattern.rs:93-109
//   compile              <- BarcodeRegex::new               pattern.rs:179-188 (+ :191-205, :137-144)
//
// The reference hands the expanded text to `regex::bytes::Regex`.  We instead lower it to a
// fixed-offset matching program (bk_program.h).  The lowering accepts the subset of the regex
// syntax whose matches have one possible length; everything else is refused with
// BK_E_UNSUPPORTED instead of being approximated.
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/barkit_b200.h"
#include "bk_internal.h"

namespace bk {

static inline bool is_word(unsigned char c) {
  return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_';
}
static inline bool in_fuzzy_alphabet(unsigned char c) {
  // pattern.rs:11  [atgcryswkmbdhvn]
  switch (c) {
    case 'a': case 't': case 'g': case 'c': case 'r': case 'y': case 's': case 'w':
    case 'k': case 'm': case 'b': case 'd': case 'h': case 'v': case 'n':
      return true;
    default:
      return false;
  }
}
static inline char ascii_upper(char c) { return (c >= 'a' && c <= 'z') ? char(c - 32) : c; }

// (?<!\[)\b[atgcryswkmbdhvn]+\b(?!\])  -- non-overlapping leftmost matches.
// A candidate start i needs: s[i] in the alphabet, s[i-1] not a word character (\b) and not '['.
// The run is greedy; any shorter end lands between two word characters, where \b cannot hold, so
// the only end that can succeed is the end of the maximal run, and it needs a non-word character
// (or the end of the pattern) that is not ']'.  After a failed start the rest of the run cannot
// start a match either (its left neighbour is a word character).
std::vector<Span> find_fuzzy_runs(const std::string& s) {
  std::vector<Span> out;
  size_t n = s.size(), i = 0;
  while (i < n) {
    unsigned char c = s[i];
    if (!in_fuzzy_alphabet(c)) { ++i; continue; }
    size_t j = i;
    while (j < n && in_fuzzy_alphabet((unsigned char)s[j])) ++j;
    bool left_ok = (i == 0) || (!is_word((unsigned char)s[i - 1]) && s[i - 1] != '[');
    bool right_ok = (j == n) || (!is_word((unsigned char)s[j]) && s[j] != ']');
    if (left_ok && right_ok) out.push_back({i, j});
    // skip the rest of this word: no position inside it has a \b on its left
    while (j < n && is_word((unsigned char)s[j])) ++j;
    i = j;
  }
  return out;
}

static const size_t kMaxExpandBytes = size_t(256) << 20;

int sequence_with_errors(const std::string& seq, size_t k, std::vector<std::string>* out,
                         std::string* err) {
  out->clear();
  size_t n = seq.size();
  std::string upper(seq);
  for (auto& ch : upper) ch = ascii_upper(ch);
  if (k == 0) { out->push_back(upper); return BK_OK; }          // pattern.rs:41-43
  if (n == 0) return BK_OK;                                      // pattern.rs:45-47
  if (k >= n) { out->push_back(std::string(n, '.')); return BK_OK; }  // pattern.rs:49-51
  if (n > 64) {  // pattern.rs:56-58: the shift amount underflows, checked_shr -> None
    *err = "Failed to choose permutation mask";
    return BK_E_PATTERN;
  }
  // pattern.rs:64-77 keeps masks with popcount n-k in ascending order; bit i set = char i kept.
  // Count first so that absurd sizes are refused instead of exhausting memory.
  long double count = 1;
  for (size_t i = 0; i < k; ++i) count = count * (long double)(n - i) / (long double)(i + 1);
  if (count * (long double)(n + 1) > (long double)kMaxExpandBytes) {
    *err = "fuzzy expansion too large";
    return BK_E_CAPACITY;
  }
  size_t keep = n - k;
  unsigned __int128 mask = (((unsigned __int128)1) << keep) - 1;
  unsigned __int128 limit = ((unsigned __int128)1) << n;
  while (mask < limit) {
    std::string s(n, '.');
    for (size_t i = 0; i < n; ++i)
      if ((mask >> i) & 1) s[i] = upper[i];
    out->push_back(s);
    unsigned __int128 c = mask & (~mask + 1);  // Gosper: next mask with the same popcount
    unsigned __int128 r = mask + c;
    mask = (((r ^ mask) >> 2) / c) | r;
  }
  return BK_OK;
}

int expand(const std::string& pattern, size_t k, std::string* out, std::vector<FuzzySpan>* spans,
           std::string* err) {
  for (unsigned char c : pattern)
    if (c >= 0x80) {
      *err = "non-ASCII pattern characters are not supported";
      return BK_E_UNSUPPORTED;
    }
  out->clear();
  if (spans) spans->clear();
  size_t last = 0;
  for (const Span& m : find_fuzzy_runs(pattern)) {
    out->append(pattern, last, m.begin - last);
    std::vector<std::string> alts;
    std::string lit = pattern.substr(m.begin, m.end - m.begin);
    int rc = sequence_with_errors(lit, k, &alts, err);
    if (rc != BK_OK) return rc;
    size_t b = out->size();
    out->push_back('(');
    for (size_t i = 0; i < alts.size(); ++i) {
      if (i) out->push_back('|');
      out->append(alts[i]);
    }
    out->push_back(')');
    if (spans) {
      FuzzySpan fs;
      fs.begin = b;
      fs.end = out->size();
      fs.lit = lit;
      for (auto& ch : fs.lit) ch = ascii_upper(ch);
      fs.k = k < lit.size() ? k : lit.size();
      spans->push_back(fs);
    }
    last = m.end;
  }
  out->append(pattern, last, std::string::npos);
  return BK_OK;
}

// ---- lowering ---------------------------------------------------------------------------------

namespace {

struct Pos {          // one byte position of a match
  uint8_t kind;       // 0 class, 1 literal
  uint8_t byte;       // literal byte
  int cls;            // class index
  int group;          // fuzzy group id (-1: strict)
};

struct Class128 {
  uint32_t w[4];
  bool operator==(const Class128& o) const { return !memcmp(w, o.w, sizeof w); }
};

struct Lowering {
  const std::string& s;                  // expanded pattern
  const std::vector<FuzzySpan>& spans;
  std::vector<Pos> pos;
  std::vector<Class128> classes;
  std::vector<size_t> group_k;           // per fuzzy group
  struct CapRec { std::string name; size_t off, len; size_t ins_before = 0; };
  std::vector<CapRec> caps;
  // variable repetition of a one-byte atom, kept as an insertion point (parametric lowering, compile_par below)
  bool allow_var = false;
  struct InsRec { size_t pos; long max_extra; bool lazy; int cls; uint32_t capmask; };
  std::vector<InsRec> ins;
  std::vector<size_t> open_caps;  // named groups the parser is inside of
  std::string err;
  int status = BK_OK;

  Lowering(const std::string& s_, const std::vector<FuzzySpan>& sp) : s(s_), spans(sp) {}

  bool fail(int st, const std::string& m) {
    if (status == BK_OK) { status = st; err = m; }
    return false;
  }
  bool unsupported(const std::string& what) {
    return fail(BK_E_UNSUPPORTED,
                "unsupported pattern construct (" + what + "): this build lowers only fixed-length "
                "patterns: ^ $ literals [classes] . {n} (groups) (?P<UMI|CB|SB>...) and lowercase "
                "fuzzy runs");
  }
  bool regex_error(const std::string& m) { return fail(BK_E_PATTERN, "Regex error: " + m); }

  const FuzzySpan* span_at(size_t i) const {
    for (const auto& f : spans)
      if (f.begin == i) return &f;
    return nullptr;
  }
  bool inside_span(size_t i) const {
    for (const auto& f : spans)
      if (i > f.begin && i < f.end) return true;
    return false;
  }

  int class_index(const Class128& c) {
    for (size_t i = 0; i < classes.size(); ++i)
      if (classes[i] == c) return (int)i;
    classes.push_back(c);
    return (int)classes.size() - 1;
  }

  // Parses a repetition operator at s[i] if present.  Returns false on error.  *rep = -1 when there is none, n for
  // `{n}`; any other operator is variable repetition: refused unless allow_var, else *rep = -2 and [*lo, *hi]
  // (*hi = -1: unbounded), *lazy.
  bool parse_repeat(size_t& i, long* rep, long* lo, long* hi, bool* lazy) {
    *rep = -1;
    *lazy = false;
    if (i >= s.size()) return true;
    char c = s[i];
    if (c == '?' || c == '*' || c == '+') {
      if (!allow_var) return unsupported(std::string("repetition operator ") + c);
      *lo = c == '+' ? 1 : 0;
      *hi = c == '?' ? 1 : -1;
      ++i;
    } else if (c == '{') {
      size_t j = i + 1;
      long v = 0;
      size_t digits = 0;
      while (j < s.size() && s[j] >= '0' && s[j] <= '9') {
        v = v * 10 + (s[j] - '0');
        if (v > BK_MAX_TOTAL_LEN) return unsupported("repetition count too large");
        ++j; ++digits;
      }
      if (digits == 0) return regex_error("repetition quantifier expects a valid decimal");
      if (j < s.size() && s[j] == ',') {
        if (!allow_var) return unsupported("{m,n} variable repetition");
        ++j;
        long w = 0;
        size_t d2 = 0;
        while (j < s.size() && s[j] >= '0' && s[j] <= '9') {
          w = w * 10 + (s[j] - '0');
          if (w > 60000) return unsupported("repetition count too large");
          ++j; ++d2;
        }
        if (j >= s.size() || s[j] != '}') return regex_error("unclosed counted repetition");
        if (d2 && w < v) return regex_error("invalid repetition count range, the start must be <= the end");
        *lo = v;
        *hi = d2 ? w : -1;
        i = j + 1;
      } else {
        if (j >= s.size() || s[j] != '}') return regex_error("unclosed counted repetition");
        i = j + 1;
        if (allow_var && i < s.size() && s[i] == '?') ++i;  // `{n}?`: laziness changes nothing
        if (i < s.size() && (s[i] == '?' || s[i] == '*' || s[i] == '+' || s[i] == '{'))
          return unsupported("stacked repetition / lazy quantifier");
        *rep = v;
        return true;
      }
    } else {
      return true;
    }
    if (i < s.size() && s[i] == '?') { *lazy = true; ++i; }
    if (i < s.size() && (s[i] == '?' || s[i] == '*' || s[i] == '+' || s[i] == '{'))
      return unsupported("stacked repetition");
    if (*hi == *lo) { *rep = *lo; return true; }  // `{n,n}`
    *rep = -2;
    return true;
  }

  bool parse_class(size_t& i, Class128* out) {
    // s[i] == '['
    size_t j = i + 1;
    bool neg = false;
    if (j < s.size() && s[j] == '^') { neg = true; ++j; }
    Class128 c{{0, 0, 0, 0}};
    auto set = [&](unsigned char b) { c.w[b >> 5] |= 1u << (b & 31); };
    bool first = true;
    for (;;) {
      if (j >= s.size()) return regex_error("unclosed character class");
      if (inside_span(j) || span_at(j)) return unsupported("lowercase run inside a character class");
      unsigned char ch = s[j];
      if (ch == ']' && !first) { ++j; break; }
      if (ch == '\\') return unsupported("escape in character class");
      if (ch == '[') return unsupported("nested / POSIX character class");
      if (ch == '&' && j + 1 < s.size() && s[j + 1] == '&') return unsupported("class set operation");
      if (ch == '~' && j + 1 < s.size() && s[j + 1] == '~') return unsupported("class set operation");
      if (ch == '-' && j + 1 < s.size() && s[j + 1] == '-') return unsupported("class set operation");
      if (ch == ']' && first) { /* literal ']' as first member */ }
      first = false;
      if (j + 2 < s.size() && s[j + 1] == '-' && s[j + 2] != ']') {
        unsigned char hi = s[j + 2];
        if (hi == '\\' || hi == '[') return unsupported("escape in character class range");
        if (hi < ch) return regex_error("invalid character class range");
        for (unsigned v = ch; v <= hi; ++v) set((unsigned char)v);
        j += 3;
      } else {
        set(ch);
        ++j;
      }
    }
    if (neg)
      for (int w = 0; w < 4; ++w) c.w[w] = ~c.w[w];
    *out = c;
    i = j;
    return true;
  }

  // Parses a sequence until ')' (when depth > 0) or end of pattern; appends positions.
  // has_named is set when a named capture was opened inside.
  bool parse_seq(size_t& i, int depth, bool* has_named) {
    while (i < s.size()) {
      char c = s[i];
      if (c == ')') {
        if (depth == 0) return regex_error("unopened group");
        return true;
      }
      size_t item_begin = pos.size();
      bool item_named = false, item_group = false;
      if (const FuzzySpan* f = span_at(i)) {
        int g = (int)group_k.size();
        group_k.push_back(f->k);
        for (unsigned char b : f->lit) pos.push_back(Pos{1, b, -1, g});
        i = f->end;
      } else if (c == '(') {
        ++i;
        std::string name;
        bool named = false;
        if (i < s.size() && s[i] == '?') {
          if (inside_span(i + 1) || span_at(i + 1)) return unsupported("lowercase run inside group flags");
          if (s.compare(i, 3, "?P<") == 0) { named = true; i += 3; }
          else if (s.compare(i, 2, "?<") == 0 && i + 2 < s.size() && s[i + 2] != '=' && s[i + 2] != '!') { named = true; i += 2; }
          else if (s.compare(i, 2, "?:") == 0) { i += 2; }
          else return unsupported("group flags / look-around");
          if (named) {
            size_t e = i;
            while (e < s.size() && s[e] != '>') {
              if (inside_span(e) || span_at(e)) return regex_error("invalid capture group character");
              unsigned char ch = s[e];
              if (!(is_word(ch) || ch == '.' || ch == '[' || ch == ']')) return regex_error("invalid capture group character");
              ++e;
            }
            if (e >= s.size()) return regex_error("unclosed capture group name");
            name = s.substr(i, e - i);
            if (name.empty()) return regex_error("empty capture group name");
            if (name[0] >= '0' && name[0] <= '9') return regex_error("invalid capture group character");
            for (auto& cr : caps)
              if (cr.name == name) return regex_error("duplicate capture group name");
            i = e + 1;
          }
        }
        size_t cap_slot = 0;
        if (named) {
          cap_slot = caps.size();
          caps.push_back({name, pos.size(), 0, ins.size()});  // pattern order = order of opening parens
          item_named = true;
          open_caps.push_back(cap_slot);
        }
        bool inner_named = false;
        item_group = true;
        if (!parse_seq(i, depth + 1, &inner_named)) return false;
        if (i >= s.size() || s[i] != ')') return regex_error("unclosed group");
        ++i;
        if (named) {
          caps[cap_slot].len = pos.size() - caps[cap_slot].off;
          open_caps.pop_back();
        }
        item_named = item_named || inner_named;
      } else if (c == '[') {
        Class128 cl;
        if (!parse_class(i, &cl)) return false;
        pos.push_back(Pos{0, 0, class_index(cl), -1});
      } else if (c == '.') {
        Class128 cl{{~0u, ~0u, ~0u, ~0u}};
        cl.w[0] &= ~(1u << '\n');  // `.` = any character except \n (non-`s` mode)
        pos.push_back(Pos{0, 0, class_index(cl), -1});
        ++i;
      } else if (c == '\\') {
        return unsupported("escape sequence");
      } else if (c == '|') {
        return unsupported("alternation");
      } else if (c == '^' || c == '$') {
        return unsupported("anchor inside the pattern");
      } else if (c == '?' || c == '*' || c == '+') {
        return regex_error("repetition operator missing expression");
      } else if (c == '{') {
        return regex_error("repetition operator missing expression");
      } else if (c == ']' || c == '}') {
        return unsupported(std::string("unbalanced ") + c);
      } else {
        pos.push_back(Pos{1, (uint8_t)c, -1, -1});
        ++i;
      }
      if (item_named && has_named) *has_named = true;
      long rep, rlo = 0, rhi = 0;
      bool rlazy = false;
      const size_t ins_before_item = ins.size();
      if (!parse_repeat(i, &rep, &rlo, &rhi, &rlazy)) return false;
      if (rep == -2) {
        // variable repetition of a one-byte atom: its minimum number of copies, then an insertion point
        if (item_group) return unsupported("variable repetition of a group");
        if (pos.size() != item_begin + 1 || pos[item_begin].group >= 0)
          return unsupported("variable repetition of a lowercase run");
        const Pos atom = pos[item_begin];
        pos.resize(item_begin);
        for (long r = 0; r < rlo; ++r) {
          pos.push_back(atom);
          if (pos.size() > BK_MAX_TOTAL_LEN) return unsupported("pattern longer than 4096 bytes");
        }
        int cls = atom.cls;
        if (atom.kind == 1) {  // a literal byte repeats as the class that holds just that byte
          Class128 one{{0, 0, 0, 0}};
          one.w[atom.byte >> 5] |= 1u << (atom.byte & 31);
          cls = class_index(one);
        }
        uint32_t mask = 0;
        for (size_t c : open_caps) mask |= 1u << c;
        if (ins.size() >= BK_MAX_INS) return unsupported("more than 4 variable repetitions");
        ins.push_back(InsRec{pos.size(), rhi < 0 ? -1 : rhi - rlo, rlazy, cls, mask});
      } else if (rep >= 0 && rep != 1) {
        if (ins.size() != ins_before_item) return unsupported("repetition of a group that contains variable repetition");
        if (item_named) return unsupported("repetition of a named capture group");
        std::vector<Pos> item(pos.begin() + item_begin, pos.end());
        pos.resize(item_begin);
        // every copy of a fuzzy group gets its own mismatch budget (the group is re-entered)
        for (long r = 0; r < rep; ++r) {
          std::vector<int> remap;
          for (Pos p : item) {
            if (p.group >= 0) {
              int g = p.group;
              if (r > 0) {
                bool found = false;
                for (size_t q = 0; q + 1 < remap.size(); q += 2)
                  if (remap[q] == g) { g = remap[q + 1]; found = true; break; }
                if (!found) {
                  remap.push_back(p.group);
                  group_k.push_back(group_k[p.group]);
                  g = (int)group_k.size() - 1;
                  remap.push_back(g);
                }
              }
              p.group = g;
            }
            pos.push_back(p);
            if (pos.size() > BK_MAX_TOTAL_LEN) return unsupported("pattern longer than 4096 bytes");
          }
        }
      }
      if (pos.size() > BK_MAX_TOTAL_LEN) return unsupported("pattern longer than 4096 bytes");
    }
    if (depth > 0) return regex_error("unclosed group");
    return true;
  }
};

}  // namespace

// allow_var: variable repetition of one-byte atoms is lowered to insertion points (a parametric program, bk_program.h)
static int compile_fixed(const std::string& pattern, size_t max_error, Program* prog, std::string* err, bool allow_var = false) {
  std::vector<FuzzySpan> spans;
  std::string expanded;
  int rc = expand(pattern, max_error, &expanded, &spans, err);
  if (rc != BK_OK) return rc;
  prog->expanded = expanded;

  // anchors: only as the very first / very last character
  size_t b = 0, e = expanded.size();
  bool a_start = false, a_end = false;
  if (b < e && expanded[b] == '^') { a_start = true; ++b; }
  if (e > b && expanded[e - 1] == '$') { a_end = true; --e; }
  std::string body = expanded.substr(b, e - b);
  for (auto& f : spans) {
    if (f.begin < b || f.end > e) { *err = "internal: fuzzy span outside body"; return BK_E_PATTERN; }
    f.begin -= b;
    f.end -= b;
  }

  Lowering lw(body, spans);
  lw.allow_var = allow_var;
  size_t i = 0;
  bool named = false;
  if (!lw.parse_seq(i, 0, &named)) {
    *err = lw.err;
    return lw.status;
  }
  // pattern.rs:191-205: names are validated in group order, then emptiness is an error
  for (auto& c : lw.caps) {
    if (c.name != "UMI" && c.name != "CB" && c.name != "SB") {
      *err = "Provided unexpected barcode capture group " + c.name;  // error.rs:13-14
      return BK_E_PATTERN;
    }
  }
  if (lw.caps.empty()) {
    *err = expanded + " capture group not found in your pattern";  // error.rs:11-12, pattern.rs:202
    return BK_E_PATTERN;
  }
  if (lw.classes.size() > BK_MAX_CLASSES) {
    *err = "unsupported pattern: more than 8 distinct character classes";
    return BK_E_UNSUPPORTED;
  }

  bk_devprog& d = prog->dev;
  memset(&d, 0, sizeof d);
  d.present = 1;
  d.anchor_start = a_start;
  d.anchor_end = a_end;
  d.total_len = (uint32_t)lw.pos.size();
  for (size_t c = 0; c < lw.classes.size(); ++c) memcpy(d.cls[c], lw.classes[c].w, 16);

  // runs: literals first (most selective), then classes; each in pattern order.  A parametric program's runs end at
  // its insertion points and are ordered by segment first.
  auto is_ins_pos = [&](size_t q) {
    for (const auto& in : lw.ins)
      if (in.pos == q) return true;
    return false;
  };
  auto segment_of = [&](size_t off) {
    size_t g = 0;
    for (const auto& in : lw.ins)
      if (in.pos <= off) ++g;
    return g;
  };
  std::vector<bk_op> lits, clss;
  size_t nlit = 0;
  for (size_t p = 0; p < lw.pos.size();) {
    size_t q = p + 1;
    const Pos& a = lw.pos[p];
    while (q < lw.pos.size() && lw.pos[q].kind == a.kind && lw.pos[q].group == a.group &&
           (a.kind == 1 || lw.pos[q].cls == a.cls) && !is_ins_pos(q))
      ++q;
    bk_op op;
    op.off = (uint16_t)p;
    op.len = (uint16_t)(q - p);
    op.kind = a.kind == 1 ? BK_OP_LIT : BK_OP_CLASS;
    op.k = 0;
    op.idx = 0;
    if (a.kind == 1) {
      nlit = (nlit + 3) & ~size_t(3);  // every literal starts on a word: the kernel compares 4 bytes at a time
      if (nlit + (q - p) > BK_MAX_LIT) { *err = "unsupported pattern: literal bytes exceed 512"; return BK_E_UNSUPPORTED; }
      op.idx = (uint16_t)nlit;
      for (size_t t = p; t < q; ++t) d.lit[nlit++] = lw.pos[t].byte;
      if (a.group >= 0) {
        size_t k = lw.group_k[a.group];
        op.k = (uint8_t)(k > 255 ? 255 : k);
        if (k > (q - p)) op.k = (uint8_t)((q - p) > 255 ? 255 : (q - p));
      }
      lits.push_back(op);
    } else {
      op.idx = (uint16_t)a.cls;
      clss.push_back(op);
    }
    p = q;
  }
  if (lits.size() + clss.size() > BK_MAX_OPS) {
    *err = "unsupported pattern: more than 48 compare runs";
    return BK_E_UNSUPPORTED;
  }
  for (size_t g = 0; g <= lw.ins.size(); ++g) {
    for (auto& o : lits)
      if (segment_of(o.off) == g) d.ops[d.nops++] = o;
    for (auto& o : clss)
      if (segment_of(o.off) == g) d.ops[d.nops++] = o;
    if (g < lw.ins.size()) {
      const auto& in = lw.ins[g];
      bk_ins& di = d.ins[g];
      di.pos = (uint16_t)in.pos;
      di.max_extra = (uint16_t)(in.max_extra < 0 || in.max_extra > 0xFFFE ? 0xFFFF : in.max_extra);
      di.cls = (uint8_t)in.cls;
      di.lazy = in.lazy ? 1 : 0;
      di.capmask = (uint8_t)in.capmask;
      di.op_end = (uint8_t)d.nops;
    }
  }
  d.nins = (uint32_t)lw.ins.size();
  if (lw.classes.size() > BK_MAX_CLASSES) {  // (a repeated literal byte adds a class)
    *err = "unsupported pattern: more than 8 distinct character classes";
    return BK_E_UNSUPPORTED;
  }
  for (size_t c = 0; c < lw.classes.size(); ++c) memcpy(d.cls[c], lw.classes[c].w, 16);

  d.ncaps = (uint32_t)lw.caps.size();  // <= 3: names are unique and drawn from {UMI, CB, SB}
  d.hdr_growth = 0;
  for (size_t c = 0; c < lw.caps.size(); ++c) {
    d.caps[c].off = (uint16_t)lw.caps[c].off;
    d.caps[c].len = (uint16_t)lw.caps[c].len;
    d.caps[c].name_len = (uint8_t)lw.caps[c].name.size();
    memcpy(d.caps[c].name, lw.caps[c].name.data(), lw.caps[c].name.size());
    d.hdr_growth += 3 + d.caps[c].name_len + 2 * d.caps[c].len;
    d.cap_ins_before[c] = (uint8_t)lw.caps[c].ins_before;
  }
  return BK_OK;
}

// ---- bounded variable repetition -------------------------------------------------
//
// `X{m,n}` and `X?` on a class, `.` or a literal byte (greedy, or lazy with a trailing `?`).  The regex
// crate resolves them by backtracking: the first quantifier tries its lengths in priority order
// (longest first when greedy), for each of them the second one does, and so on; the first combination
// under which the rest of the pattern matches wins (pattern.rs:225-230, leftmost-first).  Every
// combination is a fixed-length pattern, so the pattern is lowered to the list of those fixed-length
// programs IN THAT ORDER, and the kernel takes, among the smallest match starts, the first program
// that matches there.  Unbounded repetition (`*`, `+`, `{m,}`) and quantified groups are refused.

namespace {

struct Quant {
  size_t begin, end;  // the quantifier's text in the pattern
  long lo, hi;
  bool lazy;
};

int find_quantifiers(const std::string& p, std::vector<Quant>* out, std::string* err) {
  auto unsupported = [&](const std::string& what) {
    *err = "unsupported pattern construct (" + what + "): repetition ({n} {m,n} {m,} ? * +, greedy or lazy) applies to "
           "a class, `.` or a literal byte; a group is taken a fixed number of times";
    return BK_E_UNSUPPORTED;
  };
  bool in_class = false;
  size_t class_open = 0;
  char prev = 0;  // last character of the atom before position i (outside classes)
  // A group that contains a variable quantifier must not be repeated: `(A?C){3}` gives every copy its own
  // length in the regex crate (8 combinations), while substituting one length per quantifier would give all
  // copies the same one.  Such patterns are refused, not approximated.
  std::vector<char> open_has_var;  // one entry per open group: a variable quantifier was seen inside
  bool closed_has_var = false;     // ... for the group that was closed last (valid while prev == ')')
  for (size_t i = 0; i < p.size();) {
    const char c = p[i];
    if (in_class) {
      const bool first = i == class_open + 1 || (i == class_open + 2 && p[class_open + 1] == '^');
      if (c == ']' && !first) { in_class = false; prev = ']'; }
      ++i;
      continue;
    }
    if (c == '\\') { i += 2; prev = 'x'; continue; }  // (refused by the lowering)
    if (c == '[') { in_class = true; class_open = i; ++i; continue; }
    if (c == '(') {
      ++i;
      if (i < p.size() && p[i] == '?') {  // group flags are not quantifiers
        if (p.compare(i, 3, "?P<") == 0 || (p.compare(i, 2, "?<") == 0 && i + 2 < p.size() && p[i + 2] != '=' && p[i + 2] != '!')) {
          while (i < p.size() && p[i] != '>') ++i;
          if (i < p.size()) ++i;
        } else {
          i += 2;
        }
      }
      prev = '(';
      open_has_var.push_back(0);
      continue;
    }
    if (c == ')') {
      closed_has_var = !open_has_var.empty() && open_has_var.back();
      if (!open_has_var.empty()) open_has_var.pop_back();
      if (closed_has_var && !open_has_var.empty()) open_has_var.back() = 1;  // the enclosing group contains it too
      prev = ')';
      ++i;
      continue;
    }
    if (c == '?' || c == '{' || c == '*' || c == '+') {
      Quant q{i, i, 0, 0, false};
      if (c == '?') {
        q.lo = 0; q.hi = 1; q.end = i + 1;
      } else if (c == '*' || c == '+') {
        q.lo = c == '+' ? 1 : 0; q.hi = -1; q.end = i + 1;  // (hi = -1: unbounded)
      } else {
        size_t j = i + 1;
        long lo = 0, hi = 0;
        size_t d1 = 0, d2 = 0;
        while (j < p.size() && p[j] >= '0' && p[j] <= '9' && lo < 100000) { lo = lo * 10 + (p[j] - '0'); ++j; ++d1; }
        if (d1 == 0 || j >= p.size() || p[j] != ',') {  // {n} or malformed: the lowering's business
          if (d1 > 0 && prev == ')' && closed_has_var && lo != 1)
            return unsupported("repetition of a group that contains variable repetition");
          i = j; prev = '}'; continue;
        }
        ++j;
        while (j < p.size() && p[j] >= '0' && p[j] <= '9' && hi < 100000) { hi = hi * 10 + (p[j] - '0'); ++j; ++d2; }
        if (j >= p.size() || p[j] != '}') { *err = "Regex error: unclosed counted repetition"; return BK_E_PATTERN; }
        if (d2 == 0) hi = -1;
        else if (hi < lo) { *err = "Regex error: invalid repetition count range, the start must be <= the end"; return BK_E_PATTERN; }
        q.lo = lo; q.hi = hi; q.end = j + 1;
      }
      if (prev == ')' || prev == '(' || prev == 0 || prev == '^') {
        if (prev == ')') return unsupported("variable repetition of a group");
        *err = "Regex error: repetition operator missing expression";
        return BK_E_PATTERN;
      }
      if (q.end < p.size() && p[q.end] == '?') { q.lazy = true; ++q.end; }
      if (q.end < p.size() && (p[q.end] == '?' || p[q.end] == '*' || p[q.end] == '+' || p[q.end] == '{'))
        return unsupported("stacked repetition");
      out->push_back(q);
      for (char& g : open_has_var) g = 1;
      i = q.end;
      prev = '}';
      continue;
    }
    prev = c;
    ++i;
  }
  return BK_OK;
}

}  // namespace

// ---- alternation and repeated groups ---------------------------------------------------
//
// `(atgccat|gattaca)`, `(AC){1,3}`, `(a|b){2}`: the regex crate resolves these by backtracking as well -- an
// alternation tries its branches in order, a greedy repetition prefers one more round over stopping -- so a pattern
// that holds them is a list of fixed-length patterns in that priority order, exactly as bounded `{m,n}` on an atom is.
// The pattern is parsed into a small tree (sequence, alternation, group, repetition, atom), the tree is unfolded
// depth-first in the engine's decision order, and every unfolding is lowered by compile_fixed; parts of the tree with a
// single unfolding are re-emitted as their source text, so nothing else about the pattern changes.  Taken only when
// the pattern has a `|`, a variable repetition of a group, or a repeated group with choices inside; at most
// BK_MAX_VARIANTS unfoldings.  Refused, not approximated: named groups inside a branch or a repeated group (a group that
// may not take part in the match), repetition of a group that can match the empty string, alternation at the top level
// (anchors would bind to single branches), unbounded repetition of anything but a one-byte atom, and unbounded
// repetition together with alternatives.

namespace {

struct RNode {
  enum Type { ATOM, GROUP, CONCAT, ALT, REPEAT } type = ATOM;
  size_t b = 0, e = 0;        // source span
  std::string prefix;         // GROUP: "(" | "(?:" | "(?P<NAME>" | "(?<NAME>"
  bool named = false;         // GROUP
  long lo = 1, hi = 1;        // REPEAT (hi = -1: unbounded)
  bool lazy = false;
  std::vector<RNode> kids;
};

struct RParser {
  const std::string& p;
  std::string err;
  int status = BK_OK;
  // (general form only) the lowercase runs the reference's finder takes: each is ONE atom, so that a repetition operator
  // behind it binds to the whole run, as it does to the group the run expands to (pattern.rs:93-109)
  const std::vector<Span>* fuzzy = nullptr;
  explicit RParser(const std::string& s) : p(s) {}
  bool fail(int st, const std::string& m) { if (status == BK_OK) { status = st; err = m; } return false; }
  bool unsupported(const std::string& what) { return fail(BK_E_UNSUPPORTED, "unsupported pattern construct (" + what + ")"); }
  bool regex_error(const std::string& m) { return fail(BK_E_PATTERN, "Regex error: " + m); }

  bool parse_alt(size_t& i, int depth, RNode* out) {
    RNode alt;
    alt.type = RNode::ALT;
    alt.b = i;
    for (;;) {
      RNode seq;
      if (!parse_concat(i, depth, &seq)) return false;
      alt.kids.push_back(seq);
      if (i < p.size() && p[i] == '|') { ++i; continue; }
      break;
    }
    alt.e = i;
    if (alt.kids.size() == 1) *out = alt.kids[0];
    else *out = alt;
    return true;
  }

  bool parse_concat(size_t& i, int depth, RNode* out) {
    out->type = RNode::CONCAT;
    out->b = i;
    while (i < p.size() && p[i] != '|') {
      if (p[i] == ')') {
        if (depth == 0) return regex_error("unopened group");
        break;
      }
      RNode item;
      if (!parse_item(i, depth, &item)) return false;
      out->kids.push_back(item);
    }
    out->e = i;
    return true;
  }

  bool parse_item(size_t& i, int depth, RNode* out) {
    RNode atom;
    atom.b = i;
    const char c = p[i];
    const Span* run = nullptr;
    if (fuzzy)
      for (const auto& f : *fuzzy)
        if (f.begin == i) run = &f;
    if (run) {
      i = run->end;
    } else if (c == '(') {
      atom.type = RNode::GROUP;
      size_t j = i + 1;
      if (j < p.size() && p[j] == '?') {
        if (p.compare(j, 3, "?P<") == 0 || (p.compare(j, 2, "?<") == 0 && j + 2 < p.size() && p[j + 2] != '=' && p[j + 2] != '!')) {
          atom.named = true;
          while (j < p.size() && p[j] != '>') ++j;
          if (j >= p.size()) return regex_error("unclosed capture group name");
          ++j;
        } else if (p.compare(j, 2, "?:") == 0) {
          j += 2;
        } else {
          return unsupported("group flags / look-around");
        }
      }
      atom.prefix = p.substr(i, j - i);
      i = j;
      RNode inner;
      if (!parse_alt(i, depth + 1, &inner)) return false;
      if (i >= p.size() || p[i] != ')') return regex_error("unclosed group");
      ++i;
      atom.kids.push_back(inner);
    } else if (c == '[') {
      size_t j = i + 1;
      if (j < p.size() && p[j] == '^') ++j;
      if (j < p.size() && p[j] == ']') ++j;  // a leading ']' is a member
      while (j < p.size() && p[j] != ']') j += p[j] == '\\' ? 2 : 1;
      if (j >= p.size()) return regex_error("unclosed character class");
      i = j + 1;
    } else if (c == '\\') {
      if (i + 1 >= p.size()) return regex_error("incomplete escape sequence");
      i += 2;
    } else if (c == '?' || c == '*' || c == '+' || c == '{') {
      return regex_error("repetition operator missing expression");
    } else {
      ++i;
    }
    atom.e = i;
    // repetition suffix
    if (i < p.size() && (p[i] == '?' || p[i] == '*' || p[i] == '+' || p[i] == '{')) {
      if (c == '^' || c == '$') return regex_error("repetition operator missing expression");
      RNode rep;
      rep.type = RNode::REPEAT;
      rep.b = atom.b;
      if (p[i] == '?') { rep.lo = 0; rep.hi = 1; ++i; }
      else if (p[i] == '*') { rep.lo = 0; rep.hi = -1; ++i; }
      else if (p[i] == '+') { rep.lo = 1; rep.hi = -1; ++i; }
      else {
        size_t j = i + 1;
        long lo = 0, hi = 0;
        size_t d1 = 0, d2 = 0;
        while (j < p.size() && p[j] >= '0' && p[j] <= '9' && lo < 100000) { lo = lo * 10 + (p[j] - '0'); ++j; ++d1; }
        if (d1 == 0) return regex_error("repetition quantifier expects a valid decimal");
        if (j < p.size() && p[j] == ',') {
          ++j;
          while (j < p.size() && p[j] >= '0' && p[j] <= '9' && hi < 100000) { hi = hi * 10 + (p[j] - '0'); ++j; ++d2; }
          if (d2 == 0) hi = -1;
          else if (hi < lo) return regex_error("invalid repetition count range, the start must be <= the end");
        } else {
          hi = lo;
        }
        if (j >= p.size() || p[j] != '}') return regex_error("unclosed counted repetition");
        rep.lo = lo;
        rep.hi = hi;
        i = j + 1;
      }
      if (i < p.size() && p[i] == '?') { rep.lazy = true; ++i; }
      if (i < p.size() && (p[i] == '?' || p[i] == '*' || p[i] == '+' || p[i] == '{')) return unsupported("stacked repetition");
      rep.e = i;
      rep.kids.push_back(atom);
      *out = rep;
    } else {
      *out = atom;
    }
    return true;
  }
};

bool rnode_has(const RNode& n, bool (*pred)(const RNode&)) {
  if (pred(n)) return true;
  for (const auto& k : n.kids)
    if (rnode_has(k, pred)) return true;
  return false;
}
bool is_named_group(const RNode& n) { return n.type == RNode::GROUP && n.named; }
bool is_variable(const RNode& n) { return (n.type == RNode::REPEAT && n.lo != n.hi) || (n.type == RNode::ALT && n.kids.size() > 1); }
bool is_unbounded(const RNode& n) { return n.type == RNode::REPEAT && n.hi < 0; }
bool needs_tree(const RNode& n) {  // what find_quantifiers + substitution cannot express
  return (n.type == RNode::ALT && n.kids.size() > 1) ||
         (n.type == RNode::REPEAT && !n.kids.empty() && n.kids[0].type == RNode::GROUP &&
          (n.lo != n.hi || rnode_has(n.kids[0], is_variable)));
}
// fewest bytes a match of n can have (anchors count as atoms here; the lowering refuses them inside a pattern anyway)
long min_len(const RNode& n) {
  switch (n.type) {
    case RNode::ATOM: return 1;
    case RNode::GROUP: return min_len(n.kids[0]);
    case RNode::CONCAT: { long t = 0; for (const auto& k : n.kids) t += min_len(k); return t; }
    case RNode::ALT: { long t = -1; for (const auto& k : n.kids) { const long m = min_len(k); t = t < 0 || m < t ? m : t; } return t < 0 ? 0 : t; }
    case RNode::REPEAT: return n.lo * min_len(n.kids[0]);
  }
  return 0;
}

struct Unfolder {
  const std::string& p;
  std::string err;
  int status = BK_OK;
  explicit Unfolder(const std::string& s) : p(s) {}
  bool unsupported(const std::string& what) {
    if (status == BK_OK) { status = BK_E_UNSUPPORTED; err = "unsupported pattern construct (" + what + ")"; }
    return false;
  }
  std::string src(const RNode& n) const { return p.substr(n.b, n.e - n.b); }

  // every unfolding of n in the engine's priority order
  bool unfold(const RNode& n, std::vector<std::string>* out) {
    out->clear();
    if (!rnode_has(n, is_variable) || (n.type == RNode::REPEAT && n.hi < 0)) {  // one unfolding: the source text itself
      out->push_back(src(n));
      return true;
    }
    switch (n.type) {
      case RNode::ATOM:
        out->push_back(src(n));
        return true;
      case RNode::GROUP: {
        std::vector<std::string> inner;
        if (!unfold(n.kids[0], &inner)) return false;
        for (const auto& s : inner) out->push_back(n.prefix + s + ")");
        return true;
      }
      case RNode::CONCAT: {
        out->push_back("");
        for (const auto& k : n.kids) {
          std::vector<std::string> ks, next;
          if (!unfold(k, &ks)) return false;
          if (out->size() * ks.size() > BK_MAX_VARIANTS) return unsupported("more than " + std::to_string(BK_MAX_VARIANTS) + " combinations of alternatives and repetition counts");
          for (const auto& a : *out)      // the earlier element's choice varies slowest
            for (const auto& b : ks) next.push_back(a + b);
          out->swap(next);
        }
        return true;
      }
      case RNode::ALT: {
        for (const auto& k : n.kids) {
          if (rnode_has(k, is_named_group)) return unsupported("named capture group inside an alternative");
          std::vector<std::string> ks;
          if (!unfold(k, &ks)) return false;
          out->insert(out->end(), ks.begin(), ks.end());
          if (out->size() > BK_MAX_VARIANTS) return unsupported("more than " + std::to_string(BK_MAX_VARIANTS) + " combinations of alternatives and repetition counts");
        }
        return true;
      }
      case RNode::REPEAT: {
        const RNode& x = n.kids[0];
        if (rnode_has(x, is_named_group)) return unsupported("repetition of a named capture group");
        std::vector<std::string> xs;
        if (rnode_has(x, is_variable)) {
          // (a round that may match nothing has engine-specific rules for when the repetition ends)
          if (min_len(x) == 0) return unsupported("repetition of a group that can match the empty string");
          if (!unfold(x, &xs)) return false;
        } else {
          xs.push_back(src(x));
        }
        if (xs.size() == 1) {  // counts only: greedy from the largest, lazy from the smallest
          for (long r = n.lazy ? n.lo : n.hi; n.lazy ? r <= n.hi : r >= n.lo; r += n.lazy ? 1 : -1) {
            out->push_back(xs[0] + "{" + std::to_string(r) + "}");
            if (out->size() > BK_MAX_VARIANTS) return unsupported("more than " + std::to_string(BK_MAX_VARIANTS) + " combinations of alternatives and repetition counts");
          }
          return true;
        }
        // several unfoldings per round: depth-first in decision order -- at every round the engine prefers another
        // round (its alternatives in order) over stopping when greedy, stopping first when lazy
        return rounds(xs, n.lo, n.hi, n.lazy, 0, "", out);
      }
    }
    return true;
  }

  bool rounds(const std::vector<std::string>& xs, long lo, long hi, bool lazy, long done, const std::string& acc,
              std::vector<std::string>* out) {
    const bool may_stop = done >= lo, may_go = done < hi;
    if (lazy && may_stop) out->push_back(acc);
    if (may_go)
      for (const auto& x : xs)
        if (!rounds(xs, lo, hi, lazy, done + 1, acc + x, out)) return false;
    if (!lazy && may_stop) out->push_back(acc);
    if (out->size() > BK_MAX_VARIANTS) return unsupported("more than " + std::to_string(BK_MAX_VARIANTS) + " combinations of alternatives and repetition counts");
    return true;
  }
};

}  // namespace

// -> BK_OK and *handled = true when the pattern went through the tree form; *handled = false: not needed, use the
// quantifier substitution below
static int compile_tree(const std::string& pattern, size_t max_error, Program* prog, std::string* err, bool* handled) {
  *handled = false;
  // anchors stay outside the tree (they are atoms to it, and only legal at the very ends: compile_fixed checks)
  RParser ps(pattern);
  RNode root;
  size_t i = 0;
  if (!ps.parse_alt(i, 0, &root)) return BK_OK;  // malformed: the established path reports it in the reference's words
  if (i != pattern.size()) return BK_OK;
  if (!rnode_has(root, needs_tree)) return BK_OK;
  *handled = true;
  if (rnode_has(root, is_unbounded)) {
    *err = "unsupported pattern construct (unbounded repetition together with alternatives or repeated groups)";
    return BK_E_UNSUPPORTED;
  }
  if (root.type == RNode::ALT && root.kids.size() > 1) {
    // a top-level `a|b`: the anchors would bind to single branches (`^a|b$`), which the fixed form cannot say
    *err = "unsupported pattern construct (alternation at the top level: put it in a group)";
    return BK_E_UNSUPPORTED;
  }
  Unfolder uf(pattern);
  std::vector<std::string> forms;
  if (!uf.unfold(root, &forms)) { *err = uf.err; return uf.status; }
  {
    std::vector<FuzzySpan> spans;
    int rc = expand(pattern, max_error, &prog->expanded, &spans, err);
    if (rc != BK_OK) return rc;
  }
  const std::string expanded = prog->expanded;
  prog->variants.clear();
  for (const auto& f : forms) {
    Program one;
    int rc = compile_fixed(f, max_error, &one, err);
    if (rc != BK_OK) return rc;
    if (!prog->variants.empty()) {  // every form must rewrite the header the same way
      const bk_devprog& a = prog->variants[0];
      bool same = a.ncaps == one.dev.ncaps && a.anchor_start == one.dev.anchor_start && a.anchor_end == one.dev.anchor_end;
      for (uint32_t c = 0; same && c < a.ncaps; ++c)
        same = a.caps[c].name_len == one.dev.caps[c].name_len && !memcmp(a.caps[c].name, one.dev.caps[c].name, 3);
      if (!same) { *err = "unsupported pattern construct (alternatives with different capture groups)"; return BK_E_UNSUPPORTED; }
    }
    prog->variants.push_back(one.dev);
  }
  prog->dev = prog->variants[0];
  prog->expanded = expanded;
  return BK_OK;
}

// ---- the general form: a backtracking program (bk_program.h: bk_vmop) -----------------------------------------------
//
// For what none of the forms above can express.  The tree of compile_tree (with every lowercase run as one atom) is
// compiled the way a backtracking regex engine would run it: alternation = SPLIT first branch / rest; X{m,n} = m copies
// and n - m nested optional copies; X* = a loop whose SPLIT prefers the body (greedy) or the exit (lazy); a named group =
// SAVE start, body, SAVE end; repetition of a one-byte class is one instruction with one backtracking entry.  Refused: a
// repeated group that can match the empty string (engines differ on when such a loop ends), escapes, flags, look-around,
// and a lowercase run inside a character class (the reference's finder would rewrite the class).

namespace {

struct VmGen {
  const std::string& p;
  size_t max_error;
  const std::vector<Span>& runs;     // the lowercase runs (atoms of the tree)
  std::vector<FuzzySpan> no_spans;
  Lowering lw;                       // for its class parser and class table
  std::vector<bk_vmop> code;
  std::vector<uint8_t> lit;
  std::vector<std::pair<size_t, std::string>> names;  // named groups in pattern order: (source offset, name)
  size_t label = 0;                  // code.size() when a jump target was last taken: no literal merges across it
  std::string err;
  int status = BK_OK;
  VmGen(const std::string& s, size_t k, const std::vector<Span>& r) : p(s), max_error(k), runs(r), lw(s, no_spans) {}

  bool fail(int st, const std::string& m) { if (status == BK_OK) { status = st; err = m; } return false; }
  bool unsupported(const std::string& what) { return fail(BK_E_UNSUPPORTED, "unsupported pattern construct (" + what + ")"); }
  size_t here() { label = code.size(); return label; }
  size_t emit(uint8_t op, uint16_t a = 0, uint16_t b = 0, uint8_t k = 0, uint16_t c = 0) {
    bk_vmop o;
    o.op = op; o.k = k; o.a = a; o.b = b; o.c = c;
    code.push_back(o);
    return code.size() - 1;
  }
  bool room() { return code.size() < BK_VM_MAX ? true : unsupported("pattern needs more than " + std::to_string(BK_VM_MAX) + " instructions"); }
  const Span* run_of(const RNode& n) const {
    if (n.type != RNode::ATOM) return nullptr;
    for (const auto& f : runs)
      if (f.begin == n.b && f.end == n.e) return &f;
    return nullptr;
  }
  bool add_lit(const std::string& bytes, size_t k) {
    if (lit.size() + bytes.size() > BK_MAX_LIT) return unsupported("literal bytes exceed " + std::to_string(BK_MAX_LIT));
    const size_t at = lit.size();
    lit.insert(lit.end(), bytes.begin(), bytes.end());
    emit(BK_VM_LIT, (uint16_t)at, (uint16_t)bytes.size(), (uint8_t)(k > 255 ? 255 : k));
    return room();
  }
  int dot_class() {
    Class128 cl{{~0u, ~0u, ~0u, ~0u}};
    cl.w[0] &= ~(1u << '\n');
    return lw.class_index(cl);
  }
  // class index of a one-byte class atom (`.` or `[...]`), -1 if n is something else
  int byte_class(const RNode& n, bool* ok) {
    *ok = true;
    if (n.type != RNode::ATOM || run_of(n)) return -1;
    if (p[n.b] == '.' && n.e == n.b + 1) return dot_class();
    if (p[n.b] == '[') {
      size_t i = n.b;
      Class128 cl;
      if (!lw.parse_class(i, &cl) || i != n.e) {
        *ok = fail(lw.status != BK_OK ? lw.status : BK_E_UNSUPPORTED, lw.err.empty() ? "unsupported character class" : lw.err);
        return -1;
      }
      return lw.class_index(cl);
    }
    return -1;
  }

  // the named groups, numbered by where they open (Regex::capture_names, pattern.rs:191-205)
  bool collect_names(const RNode& n) {
    if (n.type == RNode::GROUP && n.named) {
      const size_t lt = n.prefix.find('<');
      const std::string name = n.prefix.substr(lt + 1, n.prefix.size() - lt - 2);
      if (name.empty()) return fail(BK_E_PATTERN, "Regex error: empty capture group name");
      for (char ch : name)
        if (!(is_word((unsigned char)ch) || ch == '.' || ch == '[' || ch == ']')) return fail(BK_E_PATTERN, "Regex error: invalid capture group character");
      if (name[0] >= '0' && name[0] <= '9') return fail(BK_E_PATTERN, "Regex error: invalid capture group character");
      for (const auto& nm : names)
        if (nm.second == name) return fail(BK_E_PATTERN, "Regex error: duplicate capture group name");
      names.emplace_back(n.b, name);
    }
    for (const auto& k : n.kids)
      if (!collect_names(k)) return false;
    return true;
  }
  int slot_of(const RNode& n) const {
    for (size_t c = 0; c < names.size(); ++c)
      if (names[c].first == n.b) return (int)c;
    return -1;
  }

  long min_len(const RNode& n) const {
    if (const Span* f = run_of(n)) return (long)(f->end - f->begin);
    switch (n.type) {
      case RNode::ATOM: return (p[n.b] == '^' || p[n.b] == '$') ? 0 : 1;
      case RNode::GROUP: return min_len(n.kids[0]);
      case RNode::CONCAT: { long t = 0; for (const auto& k : n.kids) t += min_len(k); return t; }
      case RNode::ALT: { long t = -1; for (const auto& k : n.kids) { const long m = min_len(k); t = t < 0 || m < t ? m : t; } return t < 0 ? 0 : t; }
      case RNode::REPEAT: return n.lo * min_len(n.kids[0]);
    }
    return 0;
  }
  // most bytes a match of n can have, -1 = no bound
  long max_len(const RNode& n) const {
    if (const Span* f = run_of(n)) return (long)(f->end - f->begin);
    switch (n.type) {
      case RNode::ATOM: return (p[n.b] == '^' || p[n.b] == '$') ? 0 : 1;
      case RNode::GROUP: return max_len(n.kids[0]);
      case RNode::CONCAT: { long t = 0; for (const auto& k : n.kids) { const long m = max_len(k); if (m < 0) return -1; t += m; } return t; }
      case RNode::ALT: { long t = 0; for (const auto& k : n.kids) { const long m = max_len(k); if (m < 0) return -1; t = m > t ? m : t; } return t; }
      case RNode::REPEAT: { const long m = max_len(n.kids[0]); return m == 0 ? 0 : (n.hi < 0 || m < 0) ? -1 : n.hi * m; }
    }
    return 0;
  }

  bool gen(const RNode& n) {
    switch (n.type) {
      case RNode::ATOM: {
        if (const Span* f = run_of(n)) {  // a lowercase run: Hamming distance <= min(k, len) (pattern.rs:40-79)
          std::string up = p.substr(f->begin, f->end - f->begin);
          for (auto& ch : up) ch = ascii_upper(ch);
          if (max_error >= up.size()) { emit(BK_VM_CLS, (uint16_t)dot_class(), (uint16_t)up.size()); return room(); }
          return add_lit(up, max_error);
        }
        const char c = p[n.b];
        if (c == '^') { emit(BK_VM_BOL); return room(); }
        if (c == '$') { emit(BK_VM_EOL); return room(); }
        if (c == '\\') return unsupported("escape sequence");
        bool ok;
        const int cls = byte_class(n, &ok);
        if (!ok) return false;
        if (cls >= 0) { emit(BK_VM_CLS, (uint16_t)cls, 1); return room(); }
        if (c == ']' || c == '}') return unsupported(std::string("unbalanced ") + c);
        if ((unsigned char)c >= 0x80) return unsupported("non-ASCII literal");
        // a literal byte joins an exact literal directly in front of it (unless a jump lands between the two)
        if (code.size() > label && code.back().op == BK_VM_LIT && code.back().k == 0 &&
            (size_t)code.back().a + code.back().b == lit.size() && lit.size() < BK_MAX_LIT) {
          lit.push_back((uint8_t)c);
          code.back().b++;
          return true;
        }
        return add_lit(std::string(1, c), 0);
      }
      case RNode::GROUP: {
        const int c = n.named ? slot_of(n) : -1;
        if (c >= 0 && c < BK_MAX_CAPS) { emit(BK_VM_SAVE, (uint16_t)(2 * c)); if (!room()) return false; }
        if (!gen(n.kids[0])) return false;
        if (c >= 0 && c < BK_MAX_CAPS) { emit(BK_VM_SAVE, (uint16_t)(2 * c + 1)); if (!room()) return false; }
        return true;
      }
      case RNode::CONCAT:
        for (const auto& k : n.kids)
          if (!gen(k)) return false;
        return true;
      case RNode::ALT: {
        std::vector<size_t> exits;
        for (size_t i = 0; i < n.kids.size(); ++i) {
          const bool last = i + 1 == n.kids.size();
          size_t split = 0;
          if (!last) {
            split = emit(BK_VM_SPLIT);
            if (!room()) return false;
            code[split].a = (uint16_t)here();
          }
          if (!gen(n.kids[i])) return false;
          if (!last) {
            exits.push_back(emit(BK_VM_JMP));
            if (!room()) return false;
            code[split].b = (uint16_t)here();
          }
        }
        for (size_t e : exits) code[e].a = (uint16_t)here();
        return true;
      }
      case RNode::REPEAT: {
        const RNode& x = n.kids[0];
        bool ok;
        const int cls = byte_class(x, &ok);
        if (!ok) return false;
        if (n.lo > 60000 || n.hi > 60000) return unsupported("repetition count too large");
        if (cls >= 0) {  // a one-byte class: the required copies, then one instruction for the optional ones
          if (n.lo > 0) { emit(BK_VM_CLS, (uint16_t)cls, (uint16_t)n.lo); if (!room()) return false; }
          if (n.hi != n.lo) { emit(BK_VM_STAR, (uint16_t)cls, (uint16_t)(n.hi < 0 ? 0xFFFF : n.hi - n.lo), n.lazy ? 1 : 0); if (!room()) return false; }
          return true;
        }
        // (a round that may match nothing has engine-specific rules for when the repetition ends)
        if ((n.hi < 0 || n.hi > 1) && min_len(x) == 0) return unsupported("repetition of a group that can match the empty string");
        if (n.lo > BK_VM_MAX || (n.hi > 0 && n.hi > BK_VM_MAX)) return unsupported("pattern needs more than " + std::to_string(BK_VM_MAX) + " instructions");
        for (long r = 0; r < n.lo; ++r)
          if (!gen(x)) return false;
        if (n.hi < 0) {  // L: SPLIT body, exit | body | JMP L
          const size_t L = emit(BK_VM_SPLIT);
          if (!room()) return false;
          const size_t b = here();
          if (!gen(x)) return false;
          emit(BK_VM_JMP, (uint16_t)L);
          if (!room()) return false;
          const size_t exit = here();
          code[L].a = (uint16_t)(n.lazy ? exit : b);
          code[L].b = (uint16_t)(n.lazy ? b : exit);
        } else {
          std::vector<size_t> splits;
          for (long r = n.lo; r < n.hi; ++r) {
            const size_t sp = emit(BK_VM_SPLIT);
            if (!room()) return false;
            splits.push_back(sp);
            code[sp].a = (uint16_t)here();  // the body (swapped below when lazy)
            if (!gen(x)) return false;
          }
          const uint16_t exit = (uint16_t)here();
          for (size_t sp : splits) {
            const uint16_t b = code[sp].a;
            code[sp].a = n.lazy ? exit : b;
            code[sp].b = n.lazy ? b : exit;
          }
        }
        return true;
      }
    }
    return true;
  }
};

}  // namespace

static int compile_vm(const std::string& pattern, size_t max_error, Program* prog, std::string* err) {
  {
    std::vector<FuzzySpan> spans;
    int rc = expand(pattern, max_error, &prog->expanded, &spans, err);  // (also the reference's own errors: pattern.rs:56-58)
    if (rc != BK_OK) return rc;
  }
  const std::vector<Span> runs = find_fuzzy_runs(pattern);
  RParser ps(pattern);
  ps.fuzzy = &runs;
  RNode root;
  size_t i = 0;
  if (!ps.parse_alt(i, 0, &root)) { *err = ps.err; return ps.status; }
  if (i != pattern.size()) { *err = "Regex error: unopened group"; return BK_E_PATTERN; }
  // a lowercase run inside a character class: the reference's finder would rewrite the class itself
  {
    bool in_class = false;
    size_t open = 0;
    for (size_t q = 0; q < pattern.size(); ++q) {
      const char c = pattern[q];
      if (!in_class && c == '\\') { ++q; continue; }
      if (!in_class && c == '[') { in_class = true; open = q; continue; }
      if (in_class && c == ']' && q > open + 1 && !(q == open + 2 && pattern[open + 1] == '^')) { in_class = false; continue; }
      if (in_class)
        for (const auto& f : runs)
          if (q >= f.begin && q < f.end) { *err = "unsupported pattern construct (lowercase run inside a character class)"; return BK_E_UNSUPPORTED; }
    }
  }
  VmGen g(pattern, max_error, runs);
  if (!g.collect_names(root) || !g.gen(root)) { *err = g.err; return g.status; }
  g.emit(BK_VM_MATCH);
  if (g.code.size() > BK_VM_MAX) { *err = "unsupported pattern construct (pattern needs more than " + std::to_string(BK_VM_MAX) + " instructions)"; return BK_E_UNSUPPORTED; }
  for (const auto& nm : g.names)
    if (nm.second != "UMI" && nm.second != "CB" && nm.second != "SB") { *err = "Provided unexpected barcode capture group " + nm.second; return BK_E_PATTERN; }
  if (g.names.empty()) { *err = prog->expanded + " capture group not found in your pattern"; return BK_E_PATTERN; }
  if (g.lw.classes.size() > BK_MAX_CLASSES) { *err = "unsupported pattern: more than 8 distinct character classes"; return BK_E_UNSUPPORTED; }
  bk_devprog& d = prog->dev;
  memset(&d, 0, sizeof d);
  d.present = 1;
  d.anchor_start = g.code[0].op == BK_VM_BOL;  // (only start 0 can match; `$` stays an instruction)
  for (size_t c = 0; c < g.lw.classes.size(); ++c) memcpy(d.cls[c], g.lw.classes[c].w, 16);
  memcpy(d.lit, g.lit.data(), g.lit.size());
  d.ncaps = (uint32_t)g.names.size();
  {
    // longest span of every capture, for the output bound (bk_ctx.cu: par_growth); 0xFFFF = as long as the read
    struct Finder {
      const VmGen& g;
      void walk(const RNode& n, bk_devprog& d) const {
        if (n.type == RNode::GROUP && n.named) {
          const int c = g.slot_of(n);
          const long m = g.max_len(n.kids[0]);
          if (c >= 0 && c < BK_MAX_CAPS) d.caps[c].len = (uint16_t)(m < 0 || m > 0xFFFF ? 0xFFFF : m);
        }
        for (const auto& k : n.kids) walk(k, d);
      }
    } finder{g};
    finder.walk(root, d);
  }
  for (size_t c = 0; c < g.names.size(); ++c) {
    d.caps[c].name_len = (uint8_t)g.names[c].second.size();
    memcpy(d.caps[c].name, g.names[c].second.data(), g.names[c].second.size());
    d.hdr_growth += 3 + d.caps[c].name_len;
  }
  // the states the device matcher enters once per read (match_vm): where paths can meet again
  uint32_t marks = 0;
  for (size_t q = 0; q < g.code.size(); ++q)
    if (g.code[q].op == BK_VM_SPLIT || g.code[q].op == BK_VM_STAR || (q > 0 && g.code[q - 1].op == BK_VM_STAR)) g.code[q].c = (uint16_t)++marks;
  d.nops = marks;
  const long tmin = g.min_len(root);
  d.total_len = (uint32_t)(tmin > BK_MAX_TOTAL_LEN ? BK_MAX_TOTAL_LEN : tmin);  // shortest match: starts beyond L - total_len are not tried
  d.vm_len = (uint32_t)g.code.size();
  prog->vm = g.code;
  prog->variants.clear();
  prog->variants.push_back(d);
  return BK_OK;
}

static int compile_lowered(const std::string& pattern, size_t max_error, Program* prog, std::string* err);

// The forms in order of preference: fixed / variants / parametric (compile_lowered); what they refuse as outside their
// grammar goes to the general form, and what that refuses too is reported with the first refusal's message.
int compile(const std::string& pattern, size_t max_error, Program* prog, std::string* err) {
  prog->vm.clear();
  const int rc = compile_lowered(pattern, max_error, prog, err);
  if (rc != BK_E_UNSUPPORTED) return rc;
  Program general;
  std::string gerr;
  const int rg = compile_vm(pattern, max_error, &general, &gerr);
  if (rg == BK_OK) { *prog = general; err->clear(); return BK_OK; }
  if (rg == BK_E_PATTERN) { *err = gerr; return rg; }
  return rc;
}

static int compile_lowered(const std::string& pattern, size_t max_error, Program* prog, std::string* err) {
  {
    bool handled = false;
    const int rct = compile_tree(pattern, max_error, prog, err, &handled);
    if (handled) return rct;
    prog->variants.clear();
  }
  std::vector<Quant> qs;
  int rc = find_quantifiers(pattern, &qs, err);
  if (rc != BK_OK) return rc;
  prog->variants.clear();
  if (qs.empty()) {
    rc = compile_fixed(pattern, max_error, prog, err);
    if (rc == BK_OK) prog->variants.push_back(prog->dev);
    return rc;
  }
  // Few length combinations: one fixed-length program each (below).  Unbounded repetition (`*`, `+`, `{m,}`) or more
  // than BK_MAX_VARIANTS combinations: ONE parametric program whose insertion points the matcher fills by backtracking
  // (bk_program.h, bk_ins).
  size_t total = 1;
  bool parametric = false;
  for (const auto& q : qs) {
    if (q.hi < 0) { parametric = true; break; }
    total *= (size_t)(q.hi - q.lo + 1);
    if (total > BK_MAX_VARIANTS) { parametric = true; break; }
  }
  if (parametric) {
    rc = compile_fixed(pattern, max_error, prog, err, true);
    if (rc == BK_OK) prog->variants.push_back(prog->dev);
    return rc;
  }
  // the expanded text the reference would hand to Regex::new keeps its quantifiers (pattern.rs:93-109)
  {
    std::vector<FuzzySpan> spans;
    rc = expand(pattern, max_error, &prog->expanded, &spans, err);
    if (rc != BK_OK) return rc;
  }
  const std::string expanded = prog->expanded;
  std::vector<long> len(qs.size());
  for (size_t i = 0; i < qs.size(); ++i) len[i] = qs[i].lazy ? qs[i].lo : qs[i].hi;
  for (size_t v = 0; v < total; ++v) {
    std::string fixed;
    size_t at = 0;
    for (size_t i = 0; i < qs.size(); ++i) {
      fixed.append(pattern, at, qs[i].begin - at);
      fixed += "{" + std::to_string(len[i]) + "}";
      at = qs[i].end;
    }
    fixed.append(pattern, at, std::string::npos);
    Program one;
    rc = compile_fixed(fixed, max_error, &one, err);
    if (rc != BK_OK) return rc;
    prog->variants.push_back(one.dev);
    // next combination in backtracking order: the LAST quantifier moves first
    for (size_t i = qs.size(); i-- > 0;) {
      const long step = qs[i].lazy ? 1 : -1, last = qs[i].lazy ? qs[i].hi : qs[i].lo;
      if (len[i] != last) { len[i] += step; break; }
      len[i] = qs[i].lazy ? qs[i].lo : qs[i].hi;
    }
  }
  prog->dev = prog->variants[0];
  prog->expanded = expanded;
  return BK_OK;
}

}  // namespace bk

// ---- C ABI ---------------------------------------------------------------------------------------

static void copy_err(const std::string& m, char* err, size_t cap) {
  if (!err || cap == 0) return;
  size_t n = m.size() < cap - 1 ? m.size() : cap - 1;
  memcpy(err, m.data(), n);
  err[n] = 0;
}

extern "C" {

int bk_sequence_with_errors(const char* sequence, size_t max_error, char* out, size_t cap,
                            size_t* n_alternatives) {
  if (!sequence || !out || cap == 0) return BK_E_ARG;
  std::vector<std::string> alts;
  std::string err;
  int rc = bk::sequence_with_errors(sequence, max_error, &alts, &err);
  if (rc != BK_OK) { copy_err(err, out, cap); return rc; }
  std::string joined;
  for (size_t i = 0; i < alts.size(); ++i) {
    if (i) joined.push_back('\n');
    joined += alts[i];
  }
  if (n_alternatives) *n_alternatives = alts.size();
  if (joined.size() + 1 > cap) return BK_E_CAPACITY;
  memcpy(out, joined.c_str(), joined.size() + 1);
  return BK_OK;
}

int bk_expand(const char* pattern, size_t max_error, char* out, size_t cap) {
  if (!pattern || !out || cap == 0) return BK_E_ARG;
  std::string expanded, err;
  int rc = bk::expand(pattern, max_error, &expanded, nullptr, &err);
  if (rc != BK_OK) { copy_err(err, out, cap); return rc; }
  if (expanded.size() + 1 > cap) return BK_E_CAPACITY;
  memcpy(out, expanded.c_str(), expanded.size() + 1);
  return BK_OK;
}

bk_program* bk_compile(const char* pattern, size_t max_error, int* status, char* err, size_t errcap) {
  if (!pattern) {
    if (status) *status = BK_E_ARG;
    copy_err("null pattern", err, errcap);
    return nullptr;
  }
  bk_program* p = new (std::nothrow) bk_program();
  if (!p) { if (status) *status = BK_E_CAPACITY; return nullptr; }
  std::string e;
  int rc = bk::compile(pattern, max_error, &p->prog, &e);
  if (status) *status = rc;
  if (rc != BK_OK) {
    copy_err(e, err, errcap);
    delete p;
    return nullptr;
  }
  if (err && errcap) err[0] = 0;
  return p;
}

void bk_program_destroy(bk_program* p) { delete p; }

int bk_program_ncaps(const bk_program* p) { return p ? (int)p->prog.dev.ncaps : 0; }

const char* bk_program_cap_name(const bk_program* p, int i) {
  if (!p || i < 0 || i >= (int)p->prog.dev.ncaps) return nullptr;
  const bk_cap& c = p->prog.dev.caps[i];
  if (c.name_len == 3) return "UMI";
  return c.name[0] == 'C' ? "CB" : "SB";
}

int bk_program_nvariants(const bk_program* p) { return p ? (int)p->prog.variants.size() : 0; }

int bk_program_get_info(const bk_program* p, bk_program_info* info) {
  if (!p || !info) return BK_E_ARG;
  const bk_devprog& d = p->prog.dev;
  memset(info, 0, sizeof *info);
  info->total_len = d.total_len;
  info->anchor_start = d.anchor_start;
  info->anchor_end = d.anchor_end;
  info->nops = d.nops;
  info->ncaps = d.ncaps;
  info->hdr_growth = d.hdr_growth;
  for (uint32_t c = 0; c < d.ncaps; ++c) {
    info->cap_off[c] = d.caps[c].off;
    info->cap_len[c] = d.caps[c].len;
  }
  return BK_OK;
}

int bk_program_dump(const bk_program* p, char* out, size_t cap) {
  if (!p || !out || cap == 0) return BK_E_ARG;
  const bk_devprog& d = p->prog.dev;
  std::string s;
  char line[128];
  snprintf(line, sizeof line, "anchor %u %u\ntotal %u\n", d.anchor_start, d.anchor_end, d.total_len);
  s += line;
  for (uint32_t o = 0; o < (d.vm_len ? 0u : d.nops); ++o) {
    const bk_op& op = d.ops[o];
    if (op.kind == BK_OP_LIT) {
      snprintf(line, sizeof line, "lit %u %u %u ", op.off, op.len, op.k);
      s += line;
      for (uint32_t i = 0; i < op.len; ++i) {
        snprintf(line, sizeof line, "%02x", d.lit[op.idx + i]);
        s += line;
      }
    } else {
      snprintf(line, sizeof line, "class %u %u ", op.off, op.len);
      s += line;
      for (int b = 0; b < 16; ++b) {
        snprintf(line, sizeof line, "%02x", (d.cls[op.idx][b >> 2] >> (8 * (b & 3))) & 0xFF);
        s += line;
      }
    }
    s += "\n";
  }
  for (uint32_t i = 0; i < d.nins; ++i) {  // insertion points of a parametric program
    snprintf(line, sizeof line, "ins %u %u %u %u %u %u\n", d.ins[i].pos, d.ins[i].max_extra, d.ins[i].lazy, d.ins[i].cls,
             d.ins[i].capmask, d.ins[i].op_end);
    s += line;
  }
  if (d.vm_len) {  // the general form: classes by index, then the instructions
    static const char* const kOp[] = {"cls", "lit", "split", "jmp", "save", "bol", "eol", "match", "star"};
    for (uint32_t k = 0; k < BK_MAX_CLASSES; ++k) {
      if (!(d.cls[k][0] | d.cls[k][1] | d.cls[k][2] | d.cls[k][3])) continue;
      snprintf(line, sizeof line, "vmclass %u ", k);
      s += line;
      for (int b = 0; b < 16; ++b) {
        snprintf(line, sizeof line, "%02x", (d.cls[k][b >> 2] >> (8 * (b & 3))) & 0xFF);
        s += line;
      }
      s += "\n";
    }
    for (uint32_t i = 0; i < d.vm_len && i < p->prog.vm.size(); ++i) {
      const bk_vmop& o = p->prog.vm[i];
      snprintf(line, sizeof line, "vm %u %s %u %u %u", i, o.op < 9 ? kOp[o.op] : "?", o.k, o.a, o.b);
      s += line;
      if (o.op == BK_VM_LIT) {
        s += " ";
        for (uint32_t j = 0; j < o.b; ++j) {
          snprintf(line, sizeof line, "%02x", d.lit[o.a + j]);
          s += line;
        }
      }
      s += "\n";
    }
  }
  for (uint32_t c = 0; c < d.ncaps; ++c) {
    snprintf(line, sizeof line, "cap %.*s %u %u\n", (int)d.caps[c].name_len, d.caps[c].name, d.caps[c].off,
             d.caps[c].len);
    s += line;
  }
  if (s.size() + 1 > cap) return BK_E_CAPACITY;
  memcpy(out, s.c_str(), s.size() + 1);
  return BK_OK;
}

}  // extern "C"
