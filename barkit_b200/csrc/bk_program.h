// Lowered matching program shared by the host compiler (bk_pattern.cc) and the kernels
// (bk_extract.cu).  POD only: it is copied verbatim into __constant__ memory.
//
// A program is what `regex::bytes::Regex` is to the reference (pattern.rs:161-167), restricted to
// the grammar whose matches have a fixed length: every element sits at a fixed offset from the
// match start, so "leftmost-first" (pattern.rs:226) reduces to "smallest start at which every
// compare run passes".  A pattern with bounded variable repetition ({m,n}, ?) is a list of such programs,
// one per length combination, in the regex crate's backtracking order (bk_pattern.cc, BK_MAX_VARIANTS).
#pragma once
#include <stdint.h>

#define BK_MAX_OPS 48
#define BK_MAX_CLASSES 8
#define BK_MAX_LIT 512
#define BK_MAX_CAPS 3
#define BK_MAX_TOTAL_LEN 4096
#define BK_MAX_VARIANTS 24
#define BK_MAX_INS 4

enum : uint8_t {
  BK_OP_CLASS = 0,  // every byte of the run must be in class `idx`
  BK_OP_LIT = 1,    // Hamming(run, lit[idx..idx+len)) <= k, mismatching bytes must be valid for `.`
};

struct bk_op {
  uint16_t off;  // offset from match start
  uint16_t len;
  uint8_t kind;
  uint8_t k;     // allowed mismatches (already min(k, len)); 0 for exact literals
  uint16_t idx;  // class index or offset into lit[] (a multiple of 4)
};

struct bk_cap {
  uint16_t off;
  uint16_t len;
  uint8_t name_len;  // 2 or 3
  char name[3];      // "UMI" | "CB" | "SB"
};

// Unbounded (or widely bounded) repetition of a one-byte atom, `X*  X+  X{m,}  X{m,n}`: the program is laid out with
// every such atom taken its minimum number of times, and each insertion point says where up to max_extra further
// copies may sit.  With e_i extra copies at insertion i, everything behind it moves by e_i: an op of segment j
// (= the ops between insertions j - 1 and j) is compared at start + off + e_0 + .. + e_{j-1}.  The matcher
// (match_par, bk_match.cuh) tries the e_i in the regex crate's backtracking order.
struct bk_ins {
  uint16_t pos;        // offset in the minimal layout at which the extra copies are inserted
  uint16_t max_extra;  // hi - lo; 0xFFFF = unbounded
  uint8_t cls;         // class index of the repeated atom
  uint8_t lazy;        // fewest copies first
  uint8_t capmask;     // bit c: capture c contains the atom (its length grows with e_i)
  uint8_t op_end;      // ops [0, op_end) belong to segments 0..i
};

// The general form (bk_pattern.cc: compile_vm): whatever the fixed, variant and parametric forms cannot express --
// unbounded repetition of groups or together with alternatives, named groups that may not take part in a match,
// alternation at the top level -- is compiled to a small backtracking program, executed one lane per read by
// match_kernel (match_vm) with the regex crate's priorities: SPLIT tries `a` before `b`.
#define BK_VM_MAX 512
enum : uint8_t {
  BK_VM_CLS = 0,    // the next b bytes all in class a
  BK_VM_LIT = 1,    // the next b bytes = lit[a .. a + b) with at most k mismatches
  BK_VM_SPLIT = 2,  // continue at a; on failure at b
  BK_VM_JMP = 3,    // continue at a
  BK_VM_SAVE = 4,   // slot a (2c: capture c starts here, 2c + 1: ends here)
  BK_VM_BOL = 5,    // at the start of the read
  BK_VM_EOL = 6,    // at the end of the read
  BK_VM_MATCH = 7,
  BK_VM_STAR = 8,   // up to b (0xFFFF: any number of) further bytes of class a, as many as possible first, or as few (k = 1)
};
struct bk_vmop {
  uint8_t op;
  uint8_t k;
  uint16_t a;
  uint16_t b;
  uint16_t c;  // 1 + this instruction's index among the marked ones (SPLIT, STAR, what follows a STAR), 0: not marked
};

struct bk_devprog {
  uint32_t present;  // 0: this mate has no pattern (run.rs:233-245 `None`)
  uint32_t anchor_start;
  uint32_t anchor_end;
  uint32_t total_len;
  uint32_t nops;
  uint32_t ncaps;
  uint32_t hdr_growth;  // sum over caps of 3 + name_len + 2*len  (parse.rs:161-168)
  uint32_t nins;        // > 0: a parametric program (ops ordered by segment; total_len, caps, hdr_growth = the minimum)
  bk_op ops[BK_MAX_OPS];
  uint32_t cls[BK_MAX_CLASSES][4];  // 128-bit ASCII membership bitmaps (bytes >= 0x80 never match)
  alignas(4) uint8_t lit[BK_MAX_LIT];  // literals, each starting on a 4-byte boundary, zero padded
  bk_cap caps[BK_MAX_CAPS];
  bk_ins ins[BK_MAX_INS];
  uint8_t cap_ins_before[4];  // capture c starts behind the insertions [0, cap_ins_before[c])
  uint32_t vm_len;            // > 0: the general form -- the code itself travels separately (Program::vm); caps[c].len
                              // is then the longest span capture c can have (0xFFFF: the read's length), total_len the
                              // shortest match, nops the number of marked instructions
};

// A program whose matches have no fixed shape -- parametric or general: match_kernel leaves the spans of every match
// (KParams::xtab) and the extract launches size and assemble every record from those.
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline bool bk_free_shape(const bk_devprog& p) { return p.nins > 0 || p.vm_len > 0; }
