// `barkit` command line: flag-for-flag the clap definition of the reference (src/lib.rs:3-114,
// src/main.rs:3-38), forwarding to bk_run (= run::run).  Extensions: --gpus N (default: one GPU per
// 2 GiB of input text per mate, at most all visible devices) and --verbose.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "../../include/barkit_b200.h"

static const char* kVersion = "0.1.1-b200";

static void usage_top(FILE* f) {
  fputs(
      "Usage: barkit [OPTIONS] <COMMAND>\n\n"
      "Commands:\n"
      "  extract  Extract barcode nucleotide sequence according to a specified regex pattern\n"
      "  help     Print this message or the help of the given subcommand(s)\n\n"
      "Options:\n"
      "  -m, --max-memory <MAX_MEMORY>  Max RAM usage in megabytes\n"
      "  -t, --threads <THREADS>        The approximate number of threads to use [default: 1]\n"
      "      --quiet                    Be quiet and do not show extra information\n"
      "  -f, --force                    Overwrite output files\n"
      "  -h, --help                     Print help\n"
      "  -V, --version                  Print version\n",
      f);
}

static void usage_extract(FILE* f) {
  fputs(
      "Extract barcode nucleotide sequence according to a specified regex pattern\n\n"
      "Usage: barkit extract [OPTIONS] --fq1 <IN_FASTQ1> --out-fq1 <OUT_FASTQ1> <--pattern1 <PATTERN1>|--pattern2 "
      "<PATTERN2>>\n\n"
      "Options:\n"
      "  -1, --fq1 <IN_FASTQ1>       Input forward FASTQ file\n"
      "  -2, --fq2 <IN_FASTQ2>       Input reverse FASTQ file\n"
      "  -o, --out-fq1 <OUT_FASTQ1>  Output forward FASTQ file\n"
      "  -O, --out-fq2 <OUT_FASTQ2>  Output reverse FASTQ file\n"
      "  -p, --pattern1 <PATTERN1>   Barcode pattern of forward reads\n"
      "  -P, --pattern2 <PATTERN2>   Barcode pattern of reverse reads\n"
      "  -t, --threads <THREADS>     The approximate number of threads to use [default: 1]\n"
      "      --gz                    Compress outputs in gzip format\n"
      "      --bgz                   Compress outputs in bgzf (bgzip) format\n"
      "      --mgz                   Compress outputs in mgzip format\n"
      "      --lz4                   Compress outputs in lz4 format\n"
      "      --quiet                 Be quiet and do not show extra information\n"
      "  -r, --rc-barcodes           Searches for both barcode pattern in reverse complement\n"
      "  -f, --force                 Overwrite output files\n"
      "  -s, --skip-trimming         Skip trimming the adapter sequence from the read\n"
      "  -e, --max-error <MAX_ERROR> Max error (mismatch) between provided pattern and read sequence [default: 1]\n"
      "      --gpus <N>              GPUs to use [default: one per 2 GiB of input per mate, at most all visible] (B200 build only)\n"
      "      --verbose               Per-stage timings on stderr (B200 build only)\n"
      "  -h, --help                  Print help\n",
      f);
}

[[noreturn]] static void die_usage(const std::string& msg) {
  fprintf(stderr, "error: %s\n\nFor more information, try '--help'.\n", msg.c_str());
  exit(2);  // clap's usage-error exit code
}

static bool parse_size(const char* s, size_t* out) {
  if (!s || !*s) return false;
  char* end = nullptr;
  unsigned long long v = strtoull(s, &end, 10);
  if (*end || s[0] == '-') return false;
  *out = (size_t)v;
  return true;
}

int main(int argc, char** argv) {
  std::string fq1, fq2, out1, out2, pat1, pat2;
  bool has_fq1 = false, has_fq2 = false, has_out1 = false, has_out2 = false, has_p1 = false, has_p2 = false;
  bool gz = false, bgz = false, mgz = false, lz4 = false, rc = false, skip = false, quiet = false, force = false;
  size_t max_memory = 0, threads = 1, max_error = 1, gpus = 0;
  bool in_extract = false, verbose = false;

  std::vector<std::string> args(argv + 1, argv + argc);
  size_t i = 0;
  auto value = [&](const std::string& name, const std::string& inline_val, bool has_inline) -> std::string {
    if (has_inline) return inline_val;
    if (i + 1 >= args.size()) die_usage("a value is required for '" + name + "' but none was supplied");
    return args[++i];
  };
  for (; i < args.size(); ++i) {
    std::string a = args[i], name = a, inl;
    bool has_inl = false;
    if (a.rfind("--", 0) == 0) {
      size_t eq = a.find('=');
      if (eq != std::string::npos) { name = a.substr(0, eq); inl = a.substr(eq + 1); has_inl = true; }
    } else if (a.size() > 2 && a[0] == '-' && a[1] != '-') {
      name = a.substr(0, 2);  // -ovalue / -o=value
      inl = a.substr(2);
      if (!inl.empty() && inl[0] == '=') inl = inl.substr(1);
      has_inl = true;
      // bundled boolean short flags (-rs) : re-expand
      bool all_flags = true;
      for (char c : a.substr(1)) all_flags = all_flags && strchr("rsfhV", c) != nullptr;
      if (all_flags) {
        std::vector<std::string> exp;
        for (char c : a.substr(1)) exp.push_back(std::string("-") + c);
        args.erase(args.begin() + i);
        args.insert(args.begin() + i, exp.begin(), exp.end());
        a = name = args[i];
        has_inl = false;
      }
    }
    if (!in_extract && (a == "extract")) { in_extract = true; continue; }
    if (a == "help") { in_extract ? usage_extract(stdout) : usage_top(stdout); return 0; }
    if (name == "-h" || name == "--help") { in_extract ? usage_extract(stdout) : usage_top(stdout); return 0; }
    if (name == "-V" || name == "--version") { printf("barkit %s\n", kVersion); return 0; }
    // global options (src/lib.rs:13-23) are accepted before and after the subcommand
    if (name == "-t" || name == "--threads") {
      if (!parse_size(value(name, inl, has_inl).c_str(), &threads)) die_usage("invalid value for '--threads <THREADS>'");
      continue;
    }
    if (name == "--quiet") { quiet = true; continue; }
    if (name == "-f" || name == "--force") { force = true; continue; }
    if (name == "-m" || name == "--max-memory") {
      if (in_extract) die_usage("unexpected argument '" + name + "' found");  // not global (src/lib.rs:9-11)
      if (!parse_size(value(name, inl, has_inl).c_str(), &max_memory)) die_usage("invalid value for '--max-memory <MAX_MEMORY>'");
      continue;
    }
    if (!in_extract) die_usage("unrecognized subcommand or argument '" + a + "'");
    if (name == "-1" || name == "--fq1") { fq1 = value(name, inl, has_inl); has_fq1 = true; }
    else if (name == "-2" || name == "--fq2") { fq2 = value(name, inl, has_inl); has_fq2 = true; }
    else if (name == "-o" || name == "--out-fq1") { out1 = value(name, inl, has_inl); has_out1 = true; }
    else if (name == "-O" || name == "--out-fq2") { out2 = value(name, inl, has_inl); has_out2 = true; }
    else if (name == "-p" || name == "--pattern1") { pat1 = value(name, inl, has_inl); has_p1 = true; }
    else if (name == "-P" || name == "--pattern2") { pat2 = value(name, inl, has_inl); has_p2 = true; }
    else if (name == "--gz") gz = true;
    else if (name == "--bgz") bgz = true;
    else if (name == "--mgz") mgz = true;
    else if (name == "--lz4") lz4 = true;
    else if (name == "-r" || name == "--rc-barcodes") rc = true;
    else if (name == "-s" || name == "--skip-trimming") skip = true;
    else if (name == "-e" || name == "--max-error") {
      if (!parse_size(value(name, inl, has_inl).c_str(), &max_error)) die_usage("invalid value for '--max-error <MAX_ERROR>'");
    } else if (name == "--gpus") {
      if (!parse_size(value(name, inl, has_inl).c_str(), &gpus)) die_usage("invalid value for '--gpus <N>'");
    } else if (name == "--verbose") {
      verbose = true;
    } else {
      die_usage("unexpected argument '" + a + "' found");
    }
  }
  if (!in_extract) { usage_top(stderr); return 2; }
  if (argc <= 2 || (!has_fq1 && !has_out1 && !has_p1 && !has_p2 && !has_fq2 && !has_out2)) {
    usage_extract(stderr);  // arg_required_else_help (src/lib.rs:29)
    return 2;
  }
  // clap constraints (src/lib.rs:51-98)
  if (!has_fq1) die_usage("the following required arguments were not provided:\n  --fq1 <IN_FASTQ1>");
  if (!has_out1) die_usage("the following required arguments were not provided:\n  --out-fq1 <OUT_FASTQ1>");
  if (!has_p1 && !has_p2)
    die_usage("the following required arguments were not provided:\n  <--pattern1 <PATTERN1>|--pattern2 <PATTERN2>>");
  if (has_fq2 && !has_out2) die_usage("the following required arguments were not provided:\n  --out-fq2 <OUT_FASTQ2>");
  if (has_p2 && !has_fq2) die_usage("the following required arguments were not provided:\n  --fq2 <IN_FASTQ2>");
  if ((int)gz + (int)bgz + (int)mgz + (int)lz4 > 1)
    die_usage("the compression arguments '--gz', '--bgz', '--mgz', '--lz4' cannot be used together");

  bk_run_args ra;
  memset(&ra, 0, sizeof ra);
  ra.fq1 = fq1.c_str();
  ra.fq2 = has_fq2 ? fq2.c_str() : nullptr;
  ra.pattern1 = has_p1 ? pat1.c_str() : nullptr;
  ra.pattern2 = has_p2 ? pat2.c_str() : nullptr;
  ra.out_fq1 = out1.c_str();
  ra.out_fq2 = has_out2 ? out2.c_str() : nullptr;
  ra.max_memory_mb = max_memory;
  ra.threads = threads;
  ra.rc_barcodes = rc;
  ra.skip_trimming = skip;
  ra.max_error = max_error;
  // src/main.rs:14-19: which flag was given; bk_run applies CompressionType::select (fastq.rs:71-79, with its
  // bgz/mgz swap)
  ra.compression = gz ? 1 : bgz ? 2 : mgz ? 3 : lz4 ? 4 : 0;
  ra.quiet = quiet;
  ra.force = force;
  ra.n_gpus = (int)gpus;
  ra.verbose = verbose;
  char err[2048];
  err[0] = 0;
  int rc_run = bk_run(&ra, err, sizeof err);
  if (rc_run != BK_OK) fprintf(stderr, "%s\n", err);  // run.rs:102-118: eprintln!("{}", e); exit(1)
  // bk_run has closed its files and released its contexts; leaving without the CUDA runtime's own exit handlers
  // saves their 0.1-0.2 s
  fflush(stdout);
  fflush(stderr);
  _exit(rc_run != BK_OK ? 1 : 0);
}
