/*
 * CPU restatement ("port") of the `barkit extract` hot path in plain C  --  TEST INFRASTRUCTURE ONLY.
 *
 * Never linked into or called by the product (barkit_b200/).  Used by tests/ as a fast second
 * checker for inputs too large for the Python oracle, and by bench.py for the `cpu_baseline` /
 * `--impl reference` figures.  It is itself validated against oracle/barkit_oracle.py (the
 * semantic anchor, pinned by the reference's known-answer tests) in tests/test_oracle_c.py.
 *
 * The reference (nsyzrantsev/barkit v0.1.1) is Rust and cannot be built in this image (no cargo /
 * rustc, no network), so this file restates its algorithm:
 *   expansion of lowercase runs into k-wildcard alternations      pattern.rs:11, :40-79, :93-109
 *   a regex compiled from the EXPANDED text and run per read        pattern.rs:182, :225-230
 *     (here: a small backtracking VM with leftmost-first priority, the match semantics of the
 *      `regex` crate 1.11.1; it executes the alternation literally, it does not use the Hamming
 *      shortcut the CUDA path uses, so it is an independent check of that lowering)
 *   per-record owned copies, header rewrite, trimming               parse.rs:58-171
 *   SE drop / PE merge policy, ordered output, write_to framing     run.rs:63-80, :149-195, fastq.rs:259-263
 *
 * Parallel structure: the reference tokenises ~64 KiB windows on one thread, maps the window's
 * records with rayon, and writes serially.  Here each thread owns one contiguous slice of the batch
 * and results are concatenated in slice order: same output, ideal scaling, so as a baseline it is
 * an UPPER bound on what the reference's structure can reach on the same cores.
 *
 * Also contains the C statement of the bksyn-v1 generator (oracle/bksyn.py is normative).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ============================ pattern.rs: finder + expansion ============================ */

static int is_word(unsigned char c) {
  return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_';
}
static int fuzzy_letter(unsigned char c) { return c && strchr("atgcryswkmbdhvn", c) != NULL; }

typedef struct { char* p; size_t n, cap; } sbuf;
static void sb_put(sbuf* b, const char* s, size_t n) {
  if (b->n + n + 1 > b->cap) {
    b->cap = (b->n + n + 1) * 2;
    b->p = (char*)realloc(b->p, b->cap);
  }
  memcpy(b->p + b->n, s, n);
  b->n += n;
  b->p[b->n] = 0;
}

/* get_sequence_with_errors (pattern.rs:40-79), appended as "(alt|alt|...)" (pattern.rs:102) */
static void put_fuzzy_group(sbuf* out, const char* lit, size_t n, size_t k) {
  char up[80];
  sb_put(out, "(", 1);
  if (n > 64) n = 64; /* longer runs are PermutationMaskSize errors in the reference; not exercised */
  for (size_t i = 0; i < n; ++i) up[i] = (lit[i] >= 'a' && lit[i] <= 'z') ? lit[i] - 32 : lit[i];
  if (k == 0) {
    sb_put(out, up, n);
  } else if (n == 0) {
  } else if (k >= n) {
    for (size_t i = 0; i < n; ++i) sb_put(out, ".", 1);
  } else {
    /* masks with popcount n-k in ascending order; bit i set keeps character i (pattern.rs:64-76) */
    unsigned __int128 mask = (((unsigned __int128)1) << (n - k)) - 1, limit = ((unsigned __int128)1) << n;
    int first = 1;
    while (mask < limit) {
      char alt[80];
      for (size_t i = 0; i < n; ++i) alt[i] = ((mask >> i) & 1) ? up[i] : '.';
      if (!first) sb_put(out, "|", 1);
      first = 0;
      sb_put(out, alt, n);
      unsigned __int128 c = mask & (~mask + 1), r = mask + c;
      mask = (((r ^ mask) >> 2) / c) | r;
    }
  }
  sb_put(out, ")", 1);
}

/* get_pattern_with_errors (pattern.rs:93-109) with the finder (?<!\[)\b[atgcryswkmbdhvn]+\b(?!\]) */
static char* expand_pattern(const char* pat, size_t k) {
  sbuf out = {0, 0, 0};
  sb_put(&out, "", 0);
  size_t n = strlen(pat), i = 0, last = 0;
  while (i < n) {
    if (!fuzzy_letter((unsigned char)pat[i])) { ++i; continue; }
    size_t j = i;
    while (j < n && fuzzy_letter((unsigned char)pat[j])) ++j;
    int left = (i == 0) || (!is_word((unsigned char)pat[i - 1]) && pat[i - 1] != '[');
    int right = (j == n) || (!is_word((unsigned char)pat[j]) && pat[j] != ']');
    if (left && right) {
      sb_put(&out, pat + last, i - last);
      put_fuzzy_group(&out, pat + i, j - i, k);
      last = j;
    }
    while (j < n && is_word((unsigned char)pat[j])) ++j;
    i = j;
  }
  sb_put(&out, pat + last, n - last);
  return out.p;
}

int cpu_expand(const char* pattern, size_t max_error, char* out, size_t cap) {
  char* e = expand_pattern(pattern, max_error);
  size_t n = strlen(e);
  if (n + 1 > cap) { free(e); return -4; }
  memcpy(out, e, n + 1);
  free(e);
  return 0;
}

/* ============================ a small leftmost-first regex VM ============================ */

enum { I_CHAR, I_ANY, I_CLASS, I_SPLIT, I_JMP, I_SAVE, I_BOL, I_EOL, I_MATCH };
typedef struct { int op; int x, y; unsigned char c; uint32_t set[8]; } inst;

#define MAXCAP 16
typedef struct {
  inst* code;
  int n, cap;
  int ncap;                 /* capture groups incl. group 0 */
  char names[MAXCAP][16];   /* "" for unnamed */
  const char* src;
  size_t pos;
  int err;
} rx;

static int emit(rx* r, int op, int x, int y) {
  if (r->n == r->cap) {
    r->cap = r->cap ? r->cap * 2 : 64;
    r->code = (inst*)realloc(r->code, sizeof(inst) * r->cap);
  }
  memset(&r->code[r->n], 0, sizeof(inst));
  r->code[r->n].op = op;
  r->code[r->n].x = x;
  r->code[r->n].y = y;
  return r->n++;
}

/* AST-free recursive-descent compiler: each parse function appends code for its fragment and
 * returns the fragment's [begin,end) so that quantifiers can duplicate or wrap it. */
typedef struct { int b, e; } frag;
static frag parse_alt(rx* r, int depth);

static void insert_at(rx* r, int at, int count) { /* open a hole of `count` instructions, fix jumps */
  for (int i = 0; i < count; ++i) emit(r, I_JMP, 0, 0);
  memmove(&r->code[at + count], &r->code[at], sizeof(inst) * (r->n - count - at));
  /* code after the hole moved with its targets; code before the hole that points AT the hole
     ("continue with what follows") must keep pointing at the new first instruction */
  for (int i = 0; i < r->n; ++i) {
    if (i >= at && i < at + count) continue;
    inst* in = &r->code[i];
    int lim = i < at ? at + 1 : at;
    if (in->op == I_SPLIT || in->op == I_JMP) {
      if (in->x >= lim) in->x += count;
      if (in->op == I_SPLIT && in->y >= lim) in->y += count;
    }
  }
}

static frag copy_frag(rx* r, frag f) {
  frag g;
  g.b = r->n;
  int delta = r->n - f.b;
  for (int i = f.b; i < f.e; ++i) {
    int k = emit(r, 0, 0, 0);
    r->code[k] = r->code[i];
    if (r->code[k].op == I_SPLIT || r->code[k].op == I_JMP) {
      r->code[k].x += delta;
      if (r->code[k].op == I_SPLIT) r->code[k].y += delta;
    }
  }
  g.e = r->n;
  return g;
}

static frag parse_atom(rx* r, int depth) {
  frag f;
  f.b = r->n;
  const char* s = r->src;
  char c = s[r->pos];
  if (c == '(') {
    r->pos++;
    int cap = -1;
    if (s[r->pos] == '?') {
      if (s[r->pos + 1] == ':') { r->pos += 2; emit(r, I_JMP, r->n + 1, 0); /* anchor for insert_at */ }
      else if ((s[r->pos + 1] == 'P' && s[r->pos + 2] == '<') || (s[r->pos + 1] == '<' && s[r->pos + 2] != '=' && s[r->pos + 2] != '!')) {
        r->pos += (s[r->pos + 1] == 'P') ? 3 : 2;
        size_t e = r->pos;
        while (s[e] && s[e] != '>') ++e;
        if (!s[e] || e - r->pos >= 16 || e == r->pos || r->ncap >= MAXCAP) { r->err = 1; return f; }
        cap = r->ncap++;
        memcpy(r->names[cap], s + r->pos, e - r->pos);
        r->names[cap][e - r->pos] = 0;
        for (int q = 1; q < cap; ++q)
          if (!strcmp(r->names[q], r->names[cap])) { r->err = 1; return f; } /* duplicate name */
        r->pos = e + 1;
      } else { r->err = 2; return f; } /* flags / look-around: outside this checker */
    } else {
      if (r->ncap >= MAXCAP) { r->err = 1; return f; }
      cap = r->ncap++;
      r->names[cap][0] = 0;
    }
    if (cap >= 0) emit(r, I_SAVE, 2 * cap, 0);
    parse_alt(r, depth + 1);
    if (r->err) return f;
    if (s[r->pos] != ')') { r->err = 1; return f; }
    r->pos++;
    if (cap >= 0) emit(r, I_SAVE, 2 * cap + 1, 0);
  } else if (c == '[') {
    int k = emit(r, I_CLASS, 0, 0);
    size_t p = r->pos + 1;
    int neg = 0, first = 1;
    if (s[p] == '^') { neg = 1; ++p; }
    uint32_t set[8] = {0};
    for (;;) {
      unsigned char ch = (unsigned char)s[p];
      if (!ch) { r->err = 1; return f; }
      if (ch == ']' && !first) { ++p; break; }
      if (ch == '\\' || ch == '[') { r->err = 2; return f; }
      first = 0;
      if (s[p + 1] == '-' && s[p + 2] && s[p + 2] != ']') {
        unsigned char hi = (unsigned char)s[p + 2];
        if (hi < ch) { r->err = 1; return f; }
        for (unsigned v = ch; v <= hi; ++v) set[v >> 5] |= 1u << (v & 31);
        p += 3;
      } else {
        set[ch >> 5] |= 1u << (ch & 31);
        ++p;
      }
    }
    for (int w = 0; w < 8; ++w) r->code[k].set[w] = neg ? ~set[w] : set[w];
    /* Unicode mode: a negated class never matches a lone byte >= 0x80 (invalid UTF-8) */
    for (int w = 4; w < 8; ++w) r->code[k].set[w] = 0;
    r->pos = p;
  } else if (c == '.') {
    emit(r, I_ANY, 0, 0);
    r->pos++;
  } else if (c == '^') {
    emit(r, I_BOL, 0, 0);
    r->pos++;
  } else if (c == '$') {
    emit(r, I_EOL, 0, 0);
    r->pos++;
  } else if (c == '\\' ) {
    r->err = 2;
    return f;
  } else if (c == '*' || c == '+' || c == '?' || c == '{') {
    r->err = 1;
    return f;
  } else {
    int k = emit(r, I_CHAR, 0, 0);
    r->code[k].c = (unsigned char)c;
    r->pos++;
  }
  f.e = r->n;
  return f;
}

static frag parse_repeat(rx* r, int depth) {
  frag f = parse_atom(r, depth);
  if (r->err) return f;
  const char* s = r->src;
  for (;;) {
    char c = s[r->pos];
    long lo = -1, hi = -1;
    if (c == '*') { lo = 0; hi = -1; r->pos++; }
    else if (c == '+') { lo = 1; hi = -1; r->pos++; }
    else if (c == '?') { lo = 0; hi = 1; r->pos++; }
    else if (c == '{') {
      size_t p = r->pos + 1;
      if (!(s[p] >= '0' && s[p] <= '9')) { r->err = 1; return f; }
      lo = 0;
      while (s[p] >= '0' && s[p] <= '9') lo = lo * 10 + (s[p++] - '0');
      hi = lo;
      if (s[p] == ',') {
        ++p;
        if (s[p] == '}') hi = -1;
        else { hi = 0; while (s[p] >= '0' && s[p] <= '9') hi = hi * 10 + (s[p++] - '0'); }
      }
      if (s[p] != '}' || lo > 5000) { r->err = 1; return f; }
      r->pos = p + 1;
    } else break;
    if (s[r->pos] == '?') { r->err = 2; return f; } /* lazy quantifiers: outside this checker */
    /* expand: lo mandatory copies, then (hi-lo) optional copies or a star */
    frag unit = copy_frag(r, f);           /* pristine copy kept at the end, removed afterwards */
    int ub = unit.b, ue = unit.e;
    /* rewrite [f.b, ub) as the repetition */
    int keep_n = ub;                        /* code before the pristine copy */
    (void)keep_n;
    /* simplest correct construction: rebuild after truncating to f.b, using the pristine copy
       saved aside */
    int ulen = ue - ub;
    inst* saved = (inst*)malloc(sizeof(inst) * (ulen ? ulen : 1));
    memcpy(saved, &r->code[ub], sizeof(inst) * ulen);
    for (int i = 0; i < ulen; ++i)
      if (saved[i].op == I_SPLIT || saved[i].op == I_JMP) {
        saved[i].x -= ub;
        if (saved[i].op == I_SPLIT) saved[i].y -= ub;
      }
    r->n = f.b;
#define PUT_UNIT()                                                  \
  do {                                                              \
    int base = r->n;                                                \
    for (int i = 0; i < ulen; ++i) {                                \
      int k = emit(r, 0, 0, 0);                                     \
      r->code[k] = saved[i];                                        \
      if (saved[i].op == I_SPLIT || saved[i].op == I_JMP) {         \
        r->code[k].x += base;                                       \
        if (saved[i].op == I_SPLIT) r->code[k].y += base;           \
      }                                                             \
    }                                                               \
  } while (0)
    for (long i = 0; i < lo; ++i) PUT_UNIT();
    if (hi < 0) { /* star: L1: split L2, L3; L2: unit; jmp L1; L3: */
      int l1 = emit(r, I_SPLIT, 0, 0);
      r->code[l1].x = r->n;
      PUT_UNIT();
      emit(r, I_JMP, l1, 0);
      r->code[l1].y = r->n;
    } else {
      int nopt = (int)(hi - lo);
      int* splits = (int*)malloc(sizeof(int) * (nopt ? nopt : 1));
      for (int i = 0; i < nopt; ++i) { /* greedy optional copies, nested */
        splits[i] = emit(r, I_SPLIT, 0, 0);
        r->code[splits[i]].x = r->n;
        PUT_UNIT();
      }
      for (int i = 0; i < nopt; ++i) r->code[splits[i]].y = r->n;
      free(splits);
    }
    free(saved);
    f.e = r->n;
  }
  return f;
}

static frag parse_cat(rx* r, int depth) {
  frag f;
  f.b = r->n;
  while (r->src[r->pos] && r->src[r->pos] != '|' && r->src[r->pos] != ')') {
    parse_repeat(r, depth);
    if (r->err) break;
  }
  f.e = r->n;
  return f;
}

static frag parse_alt(rx* r, int depth) {
  frag f = parse_cat(r, depth);
  while (!r->err && r->src[r->pos] == '|') {
    r->pos++;
    /* f | rest  ==>  split L1, L2; L1: f; jmp END; L2: rest; END: */
    insert_at(r, f.b, 1);
    r->code[f.b].op = I_SPLIT;
    r->code[f.b].x = f.b + 1;
    int j = emit(r, I_JMP, 0, 0);
    r->code[f.b].y = r->n;
    frag g = parse_alt(r, depth);
    (void)g;
    r->code[j].x = r->n;
    f.e = r->n;
    return f;
  }
  if (!r->err && r->src[r->pos] == ')' && depth == 0) r->err = 1;
  f.e = r->n;
  return f;
}

static int rx_compile(rx* r, const char* src) {
  memset(r, 0, sizeof *r);
  r->src = src;
  r->ncap = 1;
  emit(r, I_SAVE, 0, 0);
  parse_alt(r, 0);
  if (!r->err && src[r->pos]) r->err = 1;
  emit(r, I_SAVE, 1, 0);
  emit(r, I_MATCH, 0, 0);
  return r->err;
}

/* recursive backtracking, thread priority = leftmost-first */
static int rx_run(const rx* r, int pc, const unsigned char* s, long i, long n, long* caps) {
  for (;;) {
    const inst* in = &r->code[pc];
    switch (in->op) {
      case I_CHAR: if (i >= n || s[i] != in->c) return 0; ++i; ++pc; break;
      case I_ANY: if (i >= n || s[i] == '\n' || s[i] >= 0x80) return 0; ++i; ++pc; break;
      case I_CLASS: if (i >= n || !((in->set[s[i] >> 5] >> (s[i] & 31)) & 1)) return 0; ++i; ++pc; break;
      case I_BOL: if (i != 0) return 0; ++pc; break;
      case I_EOL: if (i != n) return 0; ++pc; break;
      case I_JMP: pc = in->x; break;
      case I_SPLIT: if (rx_run(r, in->x, s, i, n, caps)) return 1; pc = in->y; break;
      case I_SAVE: {
        long old = caps[in->x];
        caps[in->x] = i;
        if (rx_run(r, pc + 1, s, i, n, caps)) return 1;
        caps[in->x] = old;
        return 0;
      }
      case I_MATCH: return 1;
      default: return 0;
    }
  }
}

/* Regex::captures: leftmost start, first-priority path */
static int rx_search(const rx* r, const unsigned char* s, long n, long* caps) {
  int anchored = r->n > 1 && r->code[1].op == I_BOL;
  for (long st = 0; st <= n; ++st) {
    for (int c = 0; c < 2 * r->ncap; ++c) caps[c] = -1;
    if (rx_run(r, 0, s, st, n, caps)) return 1;
    if (anchored) break;
  }
  return 0;
}

/* ============================ BarcodeRegex (pattern.rs:161-235) ============================ */

typedef struct {
  rx re;
  char* expanded;
  int ntypes;
  int type_group[3];       /* capture index of each named group, in pattern order */
  const char* type_name[3];
} barcode_regex;

/* 0 ok; 1 regex error; 2 outside this checker's syntax; 3 unexpected name; 4 no capture group */
static int barcode_regex_new(barcode_regex* b, const char* pattern, size_t k) {
  memset(b, 0, sizeof *b);
  b->expanded = expand_pattern(pattern, k);
  int rc = rx_compile(&b->re, b->expanded);
  if (rc) return rc;
  for (int c = 1; c < b->re.ncap; ++c) {
    const char* nm = b->re.names[c];
    if (!nm[0]) continue;
    if (strcmp(nm, "UMI") && strcmp(nm, "CB") && strcmp(nm, "SB")) return 3; /* pattern.rs:137-144 */
    if (b->ntypes >= 3) return 1;
    b->type_group[b->ntypes] = c;
    b->type_name[b->ntypes] = nm;
    b->ntypes++;
  }
  return b->ntypes ? 0 : 4; /* pattern.rs:201-203 */
}
static void barcode_regex_free(barcode_regex* b) {
  free(b->re.code);
  free(b->expanded);
}

/* ============================ parse.rs ============================ */

static unsigned char TRANSLATION_TABLE[256]; /* parse.rs:12-32 */
static void init_table(void) {
  memset(TRANSLATION_TABLE, 'A', 256);
  TRANSLATION_TABLE['A'] = 'T'; TRANSLATION_TABLE['T'] = 'A';
  TRANSLATION_TABLE['G'] = 'C'; TRANSLATION_TABLE['C'] = 'G';
  for (const char* p = "RYSWKMBDHVN"; *p; ++p) TRANSLATION_TABLE[(unsigned char)*p] = (unsigned char)*p;
}

void cpu_reverse_complement(const unsigned char* seq, size_t n, unsigned char* out) { /* parse.rs:173-179 */
  if (!TRANSLATION_TABLE['A']) init_table();
  for (size_t i = 0; i < n; ++i) out[i] = TRANSLATION_TABLE[seq[n - 1 - i]];
}

typedef struct { const unsigned char *head, *seq, *qual; size_t hl, sl, ql; } ref_record;
typedef struct { unsigned char *head, *seq, *qual; size_t hl, sl, ql; } owned_record;

static void owned_free(owned_record* o) { free(o->head); free(o->seq); free(o->qual); }
static unsigned char* dup_bytes(const unsigned char* p, size_t n) {
  unsigned char* q = (unsigned char*)malloc(n ? n : 1);
  memcpy(q, p, n);
  return q;
}

/* BarcodeParser::parse_barcodes (parse.rs:58-68) + create_read (:70-90); 1 = Some(record) */
static int parse_barcodes(const barcode_regex* b, int skip_trimming, int rc_barcodes, const ref_record* rec,
                          owned_record* out, long* match_start) {
  long caps[2 * MAXCAP];
  int ok = rx_search(&b->re, rec->seq, (long)rec->sl, caps);
  if (!ok && rc_barcodes) {
    unsigned char* rcseq = (unsigned char*)malloc(rec->sl ? rec->sl : 1); /* parse.rs:62 allocates too */
    cpu_reverse_complement(rec->seq, rec->sl, rcseq);
    ok = rx_search(&b->re, rcseq, (long)rec->sl, caps);
    free(rcseq);
  }
  if (match_start) *match_start = ok ? caps[0] : -1;
  if (!ok) return 0;
  /* create_read_with_new_header (parse.rs:92-116): owned copies, then one re-allocation per barcode */
  unsigned char* head = dup_bytes(rec->head, rec->hl);
  size_t hl = rec->hl;
  unsigned char* seq = dup_bytes(rec->seq, rec->sl);
  unsigned char* qual = dup_bytes(rec->qual, rec->ql);
  for (int t = 0; t < b->ntypes; ++t) {
    long s = caps[2 * b->type_group[t]], e = caps[2 * b->type_group[t] + 1];
    if (s < 0 || (size_t)e > rec->ql) { free(head); free(seq); free(qual); return 0; } /* Err -> None */
    size_t nl = strlen(b->type_name[t]);
    size_t nh = hl + nl + 2 * (size_t)(e - s) + 3;                     /* parse.rs:161-163 */
    unsigned char* h2 = (unsigned char*)malloc(nh);
    memcpy(h2, head, hl);
    size_t p = hl;
    h2[p++] = ' ';
    memcpy(h2 + p, b->type_name[t], nl); p += nl;
    h2[p++] = ':';
    memcpy(h2 + p, seq + s, (size_t)(e - s)); p += (size_t)(e - s);
    h2[p++] = ':';
    memcpy(h2 + p, qual + s, (size_t)(e - s)); p += (size_t)(e - s);
    free(head);
    head = h2;
    hl = nh;
  }
  if (!skip_trimming) { /* trim_adapters (parse.rs:138-148) */
    long s0 = caps[0], e0 = caps[1];
    if ((size_t)e0 > rec->ql) { free(head); free(seq); free(qual); return 0; }
    unsigned char* s2 = (unsigned char*)malloc(rec->sl ? rec->sl : 1);
    unsigned char* q2 = (unsigned char*)malloc(rec->ql ? rec->ql : 1);
    memcpy(s2, seq, (size_t)s0); memcpy(s2 + s0, seq + e0, rec->sl - (size_t)e0);
    memcpy(q2, qual, (size_t)s0); memcpy(q2 + s0, qual + e0, rec->ql - (size_t)e0);
    free(seq); free(qual);
    out->seq = s2; out->sl = rec->sl - (size_t)(e0 - s0);
    out->qual = q2; out->ql = rec->ql - (size_t)(e0 - s0);
  } else {
    out->seq = seq; out->sl = rec->sl;
    out->qual = qual; out->ql = rec->ql;
  }
  out->head = head;
  out->hl = hl;
  return 1;
}

/* ============================ FASTQ tokenise / frame (seq_io stand-ins) ============================ */

static const unsigned char* next_record(const unsigned char* p, const unsigned char* end, ref_record* r) {
  const unsigned char* line[4];
  size_t len[4];
  for (int l = 0; l < 4; ++l) {
    if (p > end) return NULL;
    const unsigned char* nl = p < end ? (const unsigned char*)memchr(p, '\n', (size_t)(end - p)) : NULL;
    line[l] = p;
    if (nl) { len[l] = (size_t)(nl - p); p = nl + 1; }
    else { if (l != 3) return NULL; len[l] = (size_t)(end - p); p = end; }
    if (len[l] && line[l][len[l] - 1] == '\r') len[l]--;
  }
  if (line[0][0] != '@') return NULL;
  r->head = line[0] + 1; r->hl = len[0] - 1;
  r->seq = line[1]; r->sl = len[1];
  r->qual = line[3]; r->ql = len[3];
  return p;
}

static size_t write_to(unsigned char* out, const unsigned char* h, size_t hl, const unsigned char* s, size_t sl,
                       const unsigned char* q, size_t ql) { /* fastq.rs:261 */
  size_t p = 0;
  out[p++] = '@'; memcpy(out + p, h, hl); p += hl; out[p++] = '\n';
  memcpy(out + p, s, sl); p += sl; out[p++] = '\n'; out[p++] = '+'; out[p++] = '\n';
  memcpy(out + p, q, ql); p += ql; out[p++] = '\n';
  return p;
}

/* ============================ run.rs: batch map, ideal chunked parallelism ============================ */

typedef struct {
  const barcode_regex *b1, *b2;
  int paired, skip, rc;
  const unsigned char *t1, *e1, *t2, *e2; /* this slice's text ranges */
  unsigned char *o1, *o2;                 /* slice-local output buffers */
  size_t ol1, ol2, n_in, n_out;
  int32_t *m1, *m2;                       /* optional match tables (slice-local) */
} slice_job;

static void* slice_main(void* arg) {
  slice_job* j = (slice_job*)arg;
  const unsigned char *p1 = j->t1, *p2 = j->t2;
  /* j->o1 / j->o2 point into the persistent scratch arena; 3x input + 4096 bounds any output */
  while (p1 < j->e1) {
    ref_record r1, r2;
    p1 = next_record(p1, j->e1, &r1);
    if (!p1) break;
    long ms1 = -1, ms2 = -1;
    if (!j->paired) { /* parse_se_reads run.rs:63-80 */
      owned_record o;
      if (parse_barcodes(j->b1, j->skip, j->rc, &r1, &o, &ms1)) {
        j->ol1 += write_to(j->o1 + j->ol1, o.head, o.hl, o.seq, o.sl, o.qual, o.ql);
        owned_free(&o);
        j->n_out++;
      }
    } else { /* parse_pe_reads run.rs:163-195 + get_new_reads :149-160 */
      p2 = next_record(p2, j->e2, &r2);
      if (!p2) break;
      owned_record a, b;
      int ha = j->b1 ? parse_barcodes(j->b1, j->skip, j->rc, &r1, &a, &ms1) : 0;
      int hb = j->b2 ? parse_barcodes(j->b2, j->skip, j->rc, &r2, &b, &ms2) : 0;
      if (ha || hb) {
        if (!ha) { a.head = dup_bytes(r1.head, r1.hl); a.hl = r1.hl; a.seq = dup_bytes(r1.seq, r1.sl); a.sl = r1.sl; a.qual = dup_bytes(r1.qual, r1.ql); a.ql = r1.ql; }
        if (!hb) { b.head = dup_bytes(r2.head, r2.hl); b.hl = r2.hl; b.seq = dup_bytes(r2.seq, r2.sl); b.sl = r2.sl; b.qual = dup_bytes(r2.qual, r2.ql); b.ql = r2.ql; }
        j->ol1 += write_to(j->o1 + j->ol1, a.head, a.hl, a.seq, a.sl, a.qual, a.ql);
        j->ol2 += write_to(j->o2 + j->ol2, b.head, b.hl, b.seq, b.sl, b.qual, b.ql);
        owned_free(&a); owned_free(&b);
        j->n_out++;
      }
    }
    if (j->m1) j->m1[j->n_in] = (int32_t)ms1;
    if (j->m2) j->m2[j->n_in] = (int32_t)ms2;
    j->n_in++;
  }
  return NULL;
}

/* offset of record `idx` when every `step`-th record start is wanted: a serial newline walk */
static const unsigned char* skip_records(const unsigned char* p, const unsigned char* end, size_t count) {
  for (size_t i = 0; i < count && p < end; ++i) {
    ref_record r;
    const unsigned char* q = next_record(p, end, &r);
    if (!q) return end;
    p = q;
  }
  return p;
}

/* Persistent scratch for the slices' outputs (like the reference's long-lived 128 MiB BufWriter,
 * fastq.rs:19,252-255): re-used across calls so repeated timing runs do not measure page faults. */
static unsigned char* g_arena[2];
static size_t g_arena_cap[2];
static unsigned char* arena_get(int which, size_t bytes) {
  if (g_arena_cap[which] < bytes) {
    free(g_arena[which]);
    g_arena[which] = (unsigned char*)malloc(bytes);
    g_arena_cap[which] = bytes;
  }
  return g_arena[which];
}

static size_t count_records(const unsigned char* p, const unsigned char* end) {
  size_t n = 0;
  while (p < end) {
    ref_record r;
    const unsigned char* q = next_record(p, end, &r);
    if (!q) break;
    p = q;
    ++n;
  }
  return n;
}

/*
 * flags: 1 paired, 2 skip_trimming, 4 rc_barcodes.  Returns 0, or 10+code for pattern errors,
 * -4 when an output buffer is too small.  match1/match2 nullable (n_in int32 each).
 */
int cpu_extract(const char* pattern1, const char* pattern2, size_t max_error, int flags, const unsigned char* fq1,
                size_t len1, const unsigned char* fq2, size_t len2, unsigned char* out1, size_t cap1, size_t* olen1,
                unsigned char* out2, size_t cap2, size_t* olen2, int threads, size_t* n_in, size_t* n_out,
                int32_t* match1, int32_t* match2, const uint32_t* off1, const uint32_t* off2, size_t nrec) {
  init_table();
  barcode_regex b1, b2;
  int have1 = pattern1 != NULL, have2 = pattern2 != NULL, rc;
  if (have1 && (rc = barcode_regex_new(&b1, pattern1, max_error))) return 10 + rc;
  if (have2 && (rc = barcode_regex_new(&b2, pattern2, max_error))) return 10 + rc;
  int paired = flags & 1;
  if (threads < 1) threads = 1;
  /* slice the batch at record boundaries: equal record counts per slice (both mates) */
  /* off1/off2 (nrec+1 record offsets, the same pre-tokenised batch the CUDA path receives) make
     slicing O(1); without them the records are counted by a serial walk */
  size_t total;
  if (off1) {
    total = nrec;
  } else {
    total = count_records(fq1, fq1 + len1);
    if (paired) { size_t t2 = count_records(fq2, fq2 + len2); if (t2 < total) total = t2; } /* zip truncates */
  }
  if ((size_t)threads > total) threads = total ? (int)total : 1;
  slice_job* jobs = (slice_job*)calloc((size_t)threads, sizeof(slice_job));
  pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
  const unsigned char *p1 = fq1, *p2 = fq2;
  size_t done = 0;
  for (int t = 0; t < threads; ++t) {
    size_t cnt = total / (size_t)threads + ((size_t)t < total % (size_t)threads ? 1 : 0);
    slice_job* j = &jobs[t];
    j->b1 = have1 ? &b1 : NULL; j->b2 = have2 ? &b2 : NULL;
    j->paired = paired; j->skip = (flags & 2) != 0; j->rc = (flags & 4) != 0;
    if (off1) {
      j->t1 = fq1 + off1[done]; j->e1 = fq1 + off1[done + cnt];
      if (paired) { j->t2 = fq2 + off2[done]; j->e2 = fq2 + off2[done + cnt]; }
    } else {
      j->t1 = p1; p1 = skip_records(p1, fq1 + len1, cnt); j->e1 = p1;
      if (paired) { j->t2 = p2; p2 = skip_records(p2, fq2 + len2, cnt); j->e2 = p2; }
    }
    j->m1 = match1 ? match1 + done : NULL;
    j->m2 = match2 ? match2 + done : NULL;
    done += cnt;
  }
  {
    size_t need1 = 0, need2 = 0;
    for (int t = 0; t < threads; ++t) {
      need1 += (size_t)(jobs[t].e1 - jobs[t].t1) * 3 + 4096;
      if (paired) need2 += (size_t)(jobs[t].e2 - jobs[t].t2) * 3 + 4096;
    }
    unsigned char* a1 = arena_get(0, need1);
    unsigned char* a2 = paired ? arena_get(1, need2) : NULL;
    size_t b1 = 0, b2 = 0;
    for (int t = 0; t < threads; ++t) {
      jobs[t].o1 = a1 + b1;
      b1 += (size_t)(jobs[t].e1 - jobs[t].t1) * 3 + 4096;
      if (paired) { jobs[t].o2 = a2 + b2; b2 += (size_t)(jobs[t].e2 - jobs[t].t2) * 3 + 4096; }
    }
  }
  for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, slice_main, &jobs[t]);
  size_t o1 = 0, o2 = 0, ni = 0, no = 0;
  int status = 0;
  for (int t = 0; t < threads; ++t) {
    pthread_join(th[t], NULL);
    slice_job* j = &jobs[t];
    if (o1 + j->ol1 > cap1 || (paired && o2 + j->ol2 > cap2)) status = -4;
    if (!status) {
      memcpy(out1 + o1, j->o1, j->ol1);
      if (paired) memcpy(out2 + o2, j->o2, j->ol2);
    }
    o1 += j->ol1; o2 += j->ol2; ni += j->n_in; no += j->n_out;
  }
  free(jobs); free(th);
  if (have1) barcode_regex_free(&b1);
  if (have2) barcode_regex_free(&b2);
  if (olen1) *olen1 = o1;
  if (olen2) *olen2 = o2;
  if (n_in) *n_in = ni;
  if (n_out) *n_out = no;
  return status;
}

/* timing variant for bench.py: the parallel map only (slicing done before the clock starts would
 * hide the reference's serial tokeniser; we keep the record walk inside, as the reference does) */

/* ============================ bksyn-v1 (oracle/bksyn.py is normative) ============================ */

static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

typedef struct {
  uint64_t seed, first; size_t a, b; int mate; uint32_t L; const char* lit; uint32_t ll, off, jitter;
  unsigned char* out;
} gen_job;

static void* gen_main(void* arg) {
  gen_job* g = (gen_job*)arg;
  const uint32_t L = g->L;
  const size_t rl = 23 + 2 * (size_t)L + 6;
  for (size_t r = g->a; r < g->b; ++r) {
    unsigned char* p = g->out + r * rl;
    uint64_t idx = g->first + r, ctr = 8 * idx + 4 * (uint64_t)(g->mate - 1);
    uint64_t k0 = splitmix64(g->seed ^ splitmix64(ctr)), k1 = splitmix64(g->seed ^ splitmix64(ctr + 1));
    uint64_t k2 = splitmix64(g->seed ^ splitmix64(ctr + 2)), k3 = splitmix64(g->seed ^ splitmix64(ctr + 3));
    char head[32];
    snprintf(head, sizeof head, "bk.%012llu %d:N:0:1", (unsigned long long)idx, g->mate);
    *p++ = '@'; memcpy(p, head, 23); p += 23; *p++ = '\n';
    unsigned char* bases = p;
    uint64_t wb = 0, wn = 0, wq = 0;
    for (uint32_t i = 0; i < L; ++i) {
      if ((i & 31) == 0) wb = splitmix64(k0 + (i >> 5));
      if ((i & 7) == 0) wn = splitmix64(k2 + (i >> 3));
      unsigned char b = (unsigned char)"ACGT"[(wb >> (2 * (i & 31))) & 3];
      if (((wn >> (8 * (i & 7))) & 0xFF) == 0) b = 'N';
      bases[i] = b;
    }
    if (g->ll) {
      uint32_t u = (uint32_t)(splitmix64(k3) & 0xFF);
      if (u < 250) {
        uint32_t nsub = u < 205 ? 0 : u < 230 ? 1 : u < 243 ? 2 : 3;
        uint32_t pos0 = g->off + (uint32_t)(splitmix64(k3 + 5) % (g->jitter + 1));
        if (pos0 + g->ll <= L) {
          memcpy(bases + pos0, g->lit, g->ll);
          uint32_t used[3];
          for (uint32_t t = 0; t < nsub; ++t) {
            uint64_t w = splitmix64(k3 + 1 + t);
            uint32_t q = (uint32_t)(w % g->ll);
            for (int again = 1; again;) {
              again = 0;
              for (uint32_t v = 0; v < t; ++v)
                if (used[v] == q) { q = (q + 1) % g->ll; again = 1; }
            }
            used[t] = q;
            int orig = (int)(strchr("ACGT", g->lit[q]) - "ACGT");
            bases[pos0 + q] = (unsigned char)"ACGT"[(orig + 1 + (int)((w >> 8) % 3)) % 4];
          }
        }
      }
    }
    p += L;
    *p++ = '\n'; *p++ = '+'; *p++ = '\n';
    for (uint32_t i = 0; i < L; ++i) {
      if ((i & 7) == 0) wq = splitmix64(k1 + (i >> 3));
      p[i] = (unsigned char)(33 + ((wq >> (8 * (i & 7))) & 0xFF) % 41);
    }
    p[L] = '\n';
  }
  return NULL;
}

int cpu_bksyn_generate(uint32_t read_len, const char* lit, uint32_t lit_len, uint32_t off, uint32_t jitter, int mate,
                       uint64_t seed, uint64_t first_idx, size_t n, unsigned char* out, int threads) {
  if (threads < 1) threads = 1;
  if ((size_t)threads > n) threads = n ? (int)n : 1;
  gen_job* jobs = (gen_job*)calloc((size_t)threads, sizeof(gen_job));
  pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
  for (int t = 0; t < threads; ++t) {
    gen_job* g = &jobs[t];
    g->seed = seed; g->first = first_idx; g->mate = mate; g->L = read_len;
    g->lit = lit; g->ll = lit_len; g->off = off; g->jitter = jitter; g->out = out;
    g->a = n * (size_t)t / (size_t)threads;
    g->b = n * (size_t)(t + 1) / (size_t)threads;
    pthread_create(&th[t], NULL, gen_main, g);
  }
  for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
  free(jobs); free(th);
  return 0;
}
