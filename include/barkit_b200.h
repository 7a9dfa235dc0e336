/*
 * barkit_b200 -- C ABI of the B200-native `barkit extract` hot path.
 *
 * This header is the drop-in boundary: every entry point replaces a named item of the reference
 * crate `barkit-extract` (nsyzrantsev/barkit v0.1.1; citations are file:line into that repo).  The
 * reference is pure Rust with no FFI of its own, so these are the symbols a Rust `ffi.rs` would
 * bind (the binding is shown in INTEGRATION.md).  Plain pointers and sizes only; no C++ or torch
 * types cross this boundary; nothing here throws or aborts.
 *
 * There is NO CPU implementation of the per-read work behind this API: `bk_ctx_create` fails with
 * BK_E_CUDA when no sm_100 device is usable.
 */
#ifndef BARKIT_B200_H
#define BARKIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (error.rs:3-22 classes, plus transport errors) -------------------------- */
#define BK_OK 0
#define BK_E_PATTERN (-1)     /* Error::Regex / BarcodeCaptureGroupNotFound / UnexpectedCaptureGroupName /
                                 PermutationMaskSize: the reference would reject this pattern too */
#define BK_E_UNSUPPORTED (-2) /* valid for the reference's regex engine, outside this build's grammar */
#define BK_E_CUDA (-3)
#define BK_E_CAPACITY (-4)    /* caller buffer / batch limit too small */
#define BK_E_IO (-5)          /* Error::IO */
#define BK_E_ARG (-6)
#define BK_E_FORMAT (-7)      /* malformed FASTQ text */

/* ctx flags: run.rs:10-25 arguments `rc_barcodes`, `skip_trimming`; PAIRED = the PE branch run.rs:27 */
#define BK_FLAG_PAIRED 1u
#define BK_FLAG_SKIP_TRIM 2u
#define BK_FLAG_RC 4u
/* Library behaviour that has no counterpart in run.rs (every knob is an argument; the library reads no
 * environment variable):
 *   BK_FLAG_DEVICE_BOTH_MATES  paired mode with one pattern: send the pattern-less mate to the device as well
 *                              instead of assembling its output on the host (see bk_submit)
 *   BK_FLAG_HOST_MATE_PINNED   ... keep assembling it on the host even when the context measures the host copy
 *                              to be slower than the bus
 *   BK_FLAG_VERBOSE            launch plans and per-batch timings on stderr */
#define BK_FLAG_DEVICE_BOTH_MATES 8u
#define BK_FLAG_HOST_MATE_PINNED 16u
#define BK_FLAG_VERBOSE 32u
/*   BK_FLAG_STAGE_BOTH_MATES   paired mode with one pattern, both mates on the device: run the paired form of the
 *                              extract kernel (both mates staged through shared memory) instead of the default
 *                              two-kernel form (single-end extract of the pattern's mate + compaction of the
 *                              other); for measurements and tests, the output is identical */
#define BK_FLAG_STAGE_BOTH_MATES 64u
/*   BK_FLAG_MATCH_IN_KERNEL    a pattern that needs a search (not anchored, -r, variable repetition) is by default
 *                              matched by a pass of its own (match_kernel) that leaves a match table for the
 *                              extract launches; with this flag the search runs inside the extract kernel as in
 *                              round 1 (one launch; with two patterns: the paired kernel with both mates staged);
 *                              for measurements and tests, the output is identical */
#define BK_FLAG_MATCH_IN_KERNEL 128u
/*   BK_FLAG_SPLIT_TWO_PATTERNS two patterns of which one needs a search: after match_kernel, one single-end extract
 *                              launch per mate instead of the default single paired launch that reads the match
 *                              tables; for measurements and tests, the output is identical */
#define BK_FLAG_SPLIT_TWO_PATTERNS 256u

/* Layout guards: a binding written against this header (rust/barkit-extract/src/ffi.rs) must
 * agree with these sizes; they only ever grow at the end, together with the numbers here. */
#if defined(__cplusplus)
#define BK_STATIC_ASSERT(cond, msg) static_assert(cond, msg)
#else
#define BK_STATIC_ASSERT(cond, msg) _Static_assert(cond, msg)
#endif

typedef struct bk_program bk_program; /* replaces pattern::BarcodeRegex (pattern.rs:161-235) */
typedef struct bk_ctx bk_ctx;         /* one GPU; used by one host thread at a time          */

/* ---- pattern front end (host, one-time) --------------------------------------------------- */

/* BarcodePattern::get_sequence_with_errors (pattern.rs:40-79).  Writes the alternatives joined by
 * '\n' (NUL-terminated) and their count.  Returns BK_OK, BK_E_CAPACITY or BK_E_PATTERN. */
int bk_sequence_with_errors(const char* sequence, size_t max_error, char* out, size_t cap,
                            size_t* n_alternatives);

/* BarcodePattern::get_pattern_with_errors (pattern.rs:93-109): the expanded regex text, byte for
 * byte what the reference hands to `Regex::new` (pattern.rs:182). */
int bk_expand(const char* pattern, size_t max_error, char* out, size_t cap);

/* BarcodeRegex::new (pattern.rs:179-188): finder + expansion + lowering to a fixed-offset matching
 * program.  NULL on failure with `*status` (nullable) and `err` (nullable) filled; messages follow
 * error.rs:5-22 ("Regex error: ...", "... capture group not found in your pattern",
 * "Provided unexpected barcode capture group ..."). */
bk_program* bk_compile(const char* pattern, size_t max_error, int* status, char* err, size_t errcap);
void bk_program_destroy(bk_program* p);

/* BarcodeRegex::get_barcode_types (pattern.rs:232-234): named groups in pattern order. */
int bk_program_ncaps(const bk_program* p);
const char* bk_program_cap_name(const bk_program* p, int i); /* "UMI" | "CB" | "SB" */

/* Bounded variable repetition (`X{m,n}`, `X?` on a class, `.` or a literal; lazy forms too) lowers to one
 * fixed-length program per length combination, in the regex crate's backtracking order (pattern.rs:182,
 * :225-230); a fixed-length pattern has 1.  bk_program_get_info / bk_program_dump describe the first. */
int bk_program_nvariants(const bk_program* p);

typedef struct bk_program_info {
  uint32_t total_len;    /* bytes consumed by a match (all segments are fixed-length) */
  uint32_t anchor_start; /* pattern starts with ^ */
  uint32_t anchor_end;   /* pattern ends with $ */
  uint32_t nops;         /* compare runs in the lowered program */
  uint32_t ncaps;
  uint32_t hdr_growth;   /* bytes appended to a matched record's header (parse.rs:161-168) */
  uint32_t cap_off[3];   /* capture offsets / lengths relative to the match start */
  uint32_t cap_len[3];
} bk_program_info;
BK_STATIC_ASSERT(sizeof(bk_program_info) == 48, "bk_program_info layout");
int bk_program_get_info(const bk_program* p, bk_program_info* info);

/* Text dump of the lowered program, one line per item (debug / tests):
 *   "anchor <start> <end>" | "total <len>" | "lit <off> <len> <k> <hex bytes>" |
 *   "class <off> <len> <128-bit bitmap as 32 hex digits, byte 0 first>" | "cap <NAME> <off> <len>" */
int bk_program_dump(const bk_program* p, char* out, size_t cap);

/* ---- FASTQ text helpers (host) ------------------------------------------------------------- */

/* Record boundaries of plain FASTQ text: the part of seq_io's reader (fastq.rs:166-172) the batcher
 * needs.  Scans `text[0,len)`, writes the offset of each record's '@' to rec_off[0..n] and the end
 * of the last complete record to rec_off[n] (so rec_off has n+1 entries, n <= cap_records).
 * `final_chunk` != 0 accepts a last record without trailing newline.  `consumed` = rec_off[n].
 * `max_rec_bytes` = longest record seen.  BK_E_FORMAT on a record not starting with '@', a
 * separator not starting with '+', or seq/qual length mismatch. */
int bk_tokenize(const uint8_t* text, uint64_t len, int final_chunk, uint32_t* rec_off,
                uint32_t cap_records, uint32_t* n, uint64_t* consumed, uint32_t* max_rec_bytes);

/* ---- device context ------------------------------------------------------------------------- */

int bk_device_count(void);
/* Initialises the CUDA runtime's context on `device` (a few hundred ms the first time in a process), so that a
 * host program can overlap that with its own start-up work; bk_ctx_create does it implicitly otherwise. */
int bk_warmup(int device);

/* One context per GPU.  p1 / p2 may be NULL (run.rs:233-245: a mate without a pattern); at least
 * one must be given.  Without BK_FLAG_PAIRED, p2 must be NULL (run.rs:43). */
bk_ctx* bk_ctx_create(int device, const bk_program* p1, const bk_program* p2, uint32_t flags,
                      int* status, char* err, size_t errcap);
/* Host threads bk_wait may use for the output of a pattern-less mate (default: all cores; the reference's
 * rayon pool is sized the same way, run.rs:70,172).  Callers that run one context per GPU share the cores out. */
int bk_ctx_set_host_threads(bk_ctx* ctx, int n);
void bk_ctx_destroy(bk_ctx* ctx);
const char* bk_last_error(const bk_ctx* ctx);

/* pinned host memory for batch buffers */
void* bk_host_alloc(size_t bytes);
void bk_host_free(void* p);

/* One mate of a batch: the FASTQ text of n records, contiguous, exactly as in the file, plus the
 * n+1 record offsets from bk_tokenize.  text_len < 4 GiB.  The text allocation must extend at
 * least 16 bytes past text_len (staging loads are 16-byte granular). */
typedef struct bk_mate_in {
  const uint8_t* text;
  const uint32_t* rec_off;
  uint64_t text_len;
  uint32_t max_rec_bytes;
  uint32_t reserved;
} bk_mate_in;
BK_STATIC_ASSERT(sizeof(bk_mate_in) == 32, "bk_mate_in layout");

typedef struct bk_result {
  uint64_t out1_len; /* bytes of FASTQ text produced for -o */
  uint64_t out2_len; /* bytes produced for -O (0 in SE mode) */
  uint64_t n_in;     /* records (pairs) in */
  uint64_t n_out;    /* records (pairs) kept (run.rs:71-78, :149-160) */
  uint64_t n_match1; /* mate-1 reads whose pattern matched */
  uint64_t n_match2;
  float kernel_ms;   /* device time of the extract kernel(s), CUDA events on the ctx stream */
  uint32_t launches; /* kernels launched for this batch */
  uint64_t h2d_bytes; /* bytes copied host -> device for this batch (0 for the device-resident form) */
  uint64_t d2h_bytes; /* bytes copied device -> host */
} bk_result;
BK_STATIC_ASSERT(sizeof(bk_result) == 72, "bk_result layout");

/* The batch map + writer framing: parse_se_reads (run.rs:63-80) / parse_pe_reads (run.rs:163-195)
 * followed by FastqWriter::write_all (fastq.rs:265-271, :296-305).  HOST buffers in, HOST buffers
 * out (pinned buffers from bk_host_alloc give full PCIe rate): copies the batch to the device,
 * runs the extract kernel, copies the produced FASTQ text back.  out2 / in2 NULL in SE mode.
 * Upper bound for the output of a mate: text_len + n * hdr_growth + 6 * n (see bk_out_bound). */
int bk_extract_host(bk_ctx* ctx, uint32_t n, const bk_mate_in* in1, const bk_mate_in* in2,
                    uint8_t* out1, uint64_t cap1, uint8_t* out2, uint64_t cap2, bk_result* res);

/* A mate without a pattern stays on the host.  In paired mode with exactly one pattern the other mate's
 * records are only ever copied (the pair is kept) or dropped (run.rs:149-160), so bk_submit sends the
 * pattern's mate alone and bk_wait assembles the pattern-less mate's output from the CALLER'S INPUT
 * BUFFER and the device's match table (kept records in input order, in the writer's framing,
 * fastq.rs:259-263), on the host threads set with bk_ctx_set_host_threads.  Consequence for every
 * caller: the in1 / in2 buffers of a batch must stay valid and unchanged until bk_wait for its slot
 * returns.  BK_FLAG_DEVICE_BOTH_MATES sends both mates to the device instead.
 * bk_result.h2d_bytes / d2h_bytes report what actually crossed the bus.
 *
 * double-buffered form of the same call: slot in [0, bk_ctx_nslots) */
int bk_ctx_nslots(const bk_ctx* ctx);
int bk_submit(bk_ctx* ctx, int slot, uint32_t n, const bk_mate_in* in1, const bk_mate_in* in2);
int bk_wait(bk_ctx* ctx, int slot, uint8_t* out1, uint64_t cap1, uint8_t* out2, uint64_t cap2,
            bk_result* res);

/* Raw-text form of bk_submit (seq_io's tokenising, fastq.rs:166-172, moves to the device).
 * The caller hands over the FASTQ text of exactly n whole records per mate -- it only has to know where a record
 * boundary is, e.g. from counting line ends -- and no offsets: bk_tokenize_device's kernels make them on the
 * device, asynchronously, in front of the extract kernel.  bk_wait (same call as for bk_submit) returns BK_E_FORMAT
 * when a mate does not hold n well-formed records.  Both mates go to the device in this form.  The text need not be
 * aligned; it must stay valid until bk_wait returns.  A last record without a line end is accepted (text ends
 * with it).  max_rec_hint: an upper bound of the longest record if the caller has one, else 0 (the first batch of
 * a context then waits for the tokenizer once; later batches go by what earlier ones showed, and a batch with a
 * longer record is launched again inside bk_wait). */
int bk_submit_text(bk_ctx* ctx, int slot, uint32_t n, const uint8_t* text1, uint64_t len1, const uint8_t* text2,
                   uint64_t len2, uint32_t max_rec_hint);

uint64_t bk_out_bound(const bk_ctx* ctx, int mate, uint32_t n, uint64_t text_len);

/* ---- device-resident form (inputs and outputs stay in HBM; what `value` is timed on) -------- */

int bk_dev_alloc(bk_ctx* ctx, uint64_t bytes, void** dptr);
int bk_dev_free(bk_ctx* ctx, void* dptr);
int bk_memcpy_h2d(bk_ctx* ctx, void* dst, const void* src, uint64_t bytes);
int bk_memcpy_d2h(bk_ctx* ctx, void* dst, const void* src, uint64_t bytes);

/* Same map as bk_extract_host on device pointers (bk_mate_in fields are device pointers).
 * `d_match1/2` (nullable, n int32 each) receive the match start of every read or -1: the
 * BarcodeRegex::get_captures decision (pattern.rs:225-230) made visible for tests.
 * `iters` >= 1 repeats the launch; res->kernel_ms is the mean over iters.  Text and output pointers
 * must be 16-byte aligned (the tiles move as 16-byte bulk copies); anything from bk_dev_alloc is. */
int bk_extract_device(bk_ctx* ctx, uint32_t n, const bk_mate_in* d_in1, const bk_mate_in* d_in2,
                      uint8_t* d_out1, uint64_t cap1, uint8_t* d_out2, uint64_t cap2,
                      int32_t* d_match1, int32_t* d_match2, int iters, bk_result* res);

/* Debug: the pipeline event trace of the last device launch on slot 0 -- 16 clock64() stamps per tile
 * (tools/phases.py names them) for tiles [first_tile, first_tile + count) of a launch that had `ntiles`
 * tiles.  Fails unless the library was built with -DBK_PROFILE_PHASES. */
int bk_debug_events(bk_ctx* ctx, uint32_t ntiles, uint32_t first_tile, uint32_t count, uint64_t* out);

/* bk_tokenize on the device: the same contract for text that is already in
 * device memory (16-byte aligned, >= 16 bytes of slack behind `len`, < 4 GiB), rec_off in device
 * memory with cap_records + 1 entries.  Lets a caller ship raw FASTQ text and skip the host scan.
 * Differences from bk_tokenize, all in error cases only: after BK_E_FORMAT, max_rec_bytes covers the
 * whole scan, not just the records before the malformed one. */
int bk_tokenize_device(bk_ctx* ctx, const uint8_t* d_text, uint64_t len, int final_chunk, uint32_t* d_rec_off,
                       uint32_t cap_records, uint32_t* n_out, uint64_t* consumed, uint32_t* max_rec_bytes);

/* parse::get_reverse_complement (parse.rs:173-179, table parse.rs:12-32) for one sequence, computed on
 * the device with the same complement function the kernel's -r retry uses.  Host buffers. */
int bk_reverse_complement(bk_ctx* ctx, const uint8_t* seq, size_t n, uint8_t* out);

/* bksyn-v1 synthetic reads generated on the device (bench / tests only; not a reference feature).
 * Fixed-size records: 23 + 2*read_len + 6 bytes; writes n records of `mate` and n+1 offsets. */
typedef struct bk_syn_config {
  uint32_t read_len;
  uint32_t plant_len[2]; /* per mate; 0 = none */
  uint32_t plant_off[2];
  uint32_t plant_jitter[2];
  char plant_lit[2][32];
} bk_syn_config;
BK_STATIC_ASSERT(sizeof(bk_syn_config) == 92, "bk_syn_config layout");
int bk_generate_device(bk_ctx* ctx, const bk_syn_config* cfg, int mate, uint64_t seed,
                       uint64_t first_idx, uint32_t n, uint8_t* d_text, uint32_t* d_rec_off);

/* Measurement aid (bench.py, tools/bus_peak.py; not a reference feature): the rate of this context's GPU's host
 * link.  Copies `bytes` between pinned host memory and the device `iters` times: host -> device alone, device ->
 * host alone, then both directions at once on two streams.  GB/s (1e9 bytes); duplex = both directions summed.
 * Several contexts probing at the same time share the host's memory system, which is the point. */
int bk_bus_probe(bk_ctx* ctx, uint64_t bytes, int iters, double* h2d_gbs, double* d2h_gbs, double* duplex_gbs);
/* Measurement aid (bench.py): what the host's memory system moves when `threads` threads each copy `bytes` bytes
 * from one private buffer to another `iters` times (a plain memcpy: what the host-side assembly of a pattern-less
 * mate and, from the memory's point of view, the DMA engines do).  GB/s counts bytes read + bytes written.  The end
 * to end rate of this path is bounded by this figure over the bytes a pair moves through host memory. */
int bk_host_mem_probe(uint64_t bytes, int threads, int iters, double* gbs);

/* ---- whole-run form: run::run (run.rs:10-60), file in -> file out ---------------------------- */
typedef struct bk_run_args {
  const char* fq1;
  const char* fq2;      /* nullable */
  const char* pattern1; /* nullable */
  const char* pattern2; /* nullable */
  const char* out_fq1;
  const char* out_fq2;  /* nullable */
  size_t max_memory_mb; /* 0 = unset (-m) */
  size_t threads;       /* -t: a floor, not a cap: the host pool (file reads, line-end counts, deflate; fastq.rs:119,237,243)
                           has max(-t, cores) threads like the reference's rayon pool; the per-read work has no host threads */
  int rc_barcodes;
  int skip_trimming;
  size_t max_error;
  int compression;      /* which flag was given (src/main.rs:14-19): 0 none, 1 --gz, 2 --bgz, 3 --mgz, 4 --lz4.  bk_run applies
                           CompressionType::select (fastq.rs:71-79): --gz gzip members, --bgz Mgzip, --mgz Bgzf (the
                           reference's swap is kept), --lz4 one LZ4 frame.  Inputs are sniffed by magic bytes
                           (fastq.rs:82-99): gzip / bgzf / mgzip / lz4 are read. */
  int quiet;
  int force;
  int n_gpus;           /* 0 = automatic: one GPU per 2 GiB of input text per mate, at most all visible devices */
  int verbose;          /* != 0: per-stage timing lines on stderr */
  uint32_t ctx_flags;   /* extra BK_FLAG_* bits for the workers' contexts (BK_FLAG_DEVICE_BOTH_MATES, ...) */
} bk_run_args;
BK_STATIC_ASSERT(sizeof(void*) != 8 || sizeof(bk_run_args) == 104, "bk_run_args layout");
int bk_run(const bk_run_args* args, char* err, size_t errcap);

#ifdef __cplusplus
}
#endif
#endif /* BARKIT_B200_H */
