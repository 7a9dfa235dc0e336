// Internal C++ declarations shared by the host sources (not part of the ABI).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "bk_program.h"

namespace bk {

struct Span { size_t begin, end; };

struct FuzzySpan {   // a `( alt | alt | ... )` group produced by the expansion, in expanded text
  size_t begin, end;
  std::string lit;   // upper-cased literal
  size_t k;          // min(max_error, len)
};

struct Program {
  bk_devprog dev;        // the first (highest-priority) fixed-length program
  std::string expanded;  // what the reference would pass to Regex::new (pattern.rs:182)
  // Bounded variable repetition ({m,n}, ?) lowers to one fixed-length program per length combination, in
  // the regex crate's backtracking order; a fixed-length pattern has exactly one (== dev).
  std::vector<bk_devprog> variants;
  // the general form (bk_program.h: bk_vmop), when dev.vm_len > 0
  std::vector<bk_vmop> vm;
};

std::vector<Span> find_fuzzy_runs(const std::string& pattern);
int sequence_with_errors(const std::string& seq, size_t k, std::vector<std::string>* out,
                         std::string* err);
int expand(const std::string& pattern, size_t k, std::string* out, std::vector<FuzzySpan>* spans,
           std::string* err);
int compile(const std::string& pattern, size_t max_error, Program* prog, std::string* err);

// bk_fastq.cc: output of a mate without a pattern, assembled on the host (kept records, writer's framing)
uint64_t gather_kept(const uint8_t* text, const uint32_t* off, uint32_t n, const int32_t* keep, uint8_t* out,
                     uint64_t cap, int max_threads);

}  // namespace bk

struct bk_program {
  bk::Program prog;
};
