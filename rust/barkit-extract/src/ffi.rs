//! Raw bindings of include/barkit_b200.h.  Struct layouts are pinned by the header's BK_STATIC_ASSERTs
//! (bk_mate_in 32 B, bk_result 72 B, bk_run_args 104 B on LP64) and by the const asserts below.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct BkProgram {
    _p: [u8; 0],
}
#[repr(C)]
pub struct BkCtx {
    _p: [u8; 0],
}

#[repr(C)]
pub struct BkMateIn {
    pub text: *const u8,
    pub rec_off: *const u32,
    pub text_len: u64,
    pub max_rec_bytes: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Default, Debug, Clone, Copy)]
pub struct BkResult {
    pub out1_len: u64,
    pub out2_len: u64,
    pub n_in: u64,
    pub n_out: u64,
    pub n_match1: u64,
    pub n_match2: u64,
    pub kernel_ms: f32,
    pub launches: u32,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
}

#[repr(C)]
pub struct BkRunArgs {
    pub fq1: *const c_char,
    pub fq2: *const c_char,
    pub pattern1: *const c_char,
    pub pattern2: *const c_char,
    pub out_fq1: *const c_char,
    pub out_fq2: *const c_char,
    pub max_memory_mb: usize,
    pub threads: usize,
    pub rc_barcodes: c_int,
    pub skip_trimming: c_int,
    pub max_error: usize,
    pub compression: c_int,
    pub quiet: c_int,
    pub force: c_int,
    pub n_gpus: c_int,
    pub verbose: c_int,
    pub ctx_flags: u32,
}

const _: () = assert!(std::mem::size_of::<BkMateIn>() == 32);
const _: () = assert!(std::mem::size_of::<BkResult>() == 72);
#[cfg(target_pointer_width = "64")]
const _: () = assert!(std::mem::size_of::<BkRunArgs>() == 104);

pub const BK_OK: c_int = 0;
pub const BK_E_PATTERN: c_int = -1;
pub const BK_E_UNSUPPORTED: c_int = -2;
pub const BK_E_CUDA: c_int = -3;
pub const BK_E_CAPACITY: c_int = -4;
pub const BK_E_IO: c_int = -5;
pub const BK_E_ARG: c_int = -6;
pub const BK_E_FORMAT: c_int = -7;

pub const BK_FLAG_PAIRED: u32 = 1;
pub const BK_FLAG_SKIP_TRIM: u32 = 2;
pub const BK_FLAG_RC: u32 = 4;
pub const BK_FLAG_DEVICE_BOTH_MATES: u32 = 8;
pub const BK_FLAG_HOST_MATE_PINNED: u32 = 16;
pub const BK_FLAG_VERBOSE: u32 = 32;
pub const BK_FLAG_STAGE_BOTH_MATES: u32 = 64;
pub const BK_FLAG_MATCH_IN_KERNEL: u32 = 128;
pub const BK_FLAG_SPLIT_TWO_PATTERNS: u32 = 256;

extern "C" {
    // pattern.rs
    pub fn bk_sequence_with_errors(sequence: *const c_char, max_error: usize, out: *mut c_char, cap: usize,
                                   n_alternatives: *mut usize) -> c_int;
    pub fn bk_expand(pattern: *const c_char, max_error: usize, out: *mut c_char, cap: usize) -> c_int;
    pub fn bk_compile(pattern: *const c_char, max_error: usize, status: *mut c_int, err: *mut c_char,
                      errcap: usize) -> *mut BkProgram;
    pub fn bk_program_destroy(p: *mut BkProgram);
    pub fn bk_program_ncaps(p: *const BkProgram) -> c_int;
    pub fn bk_program_cap_name(p: *const BkProgram, i: c_int) -> *const c_char;
    pub fn bk_program_nvariants(p: *const BkProgram) -> c_int;
    // fastq.rs (record boundaries of a RecordSet's buffer)
    pub fn bk_tokenize(text: *const u8, len: u64, final_chunk: c_int, rec_off: *mut u32, cap: u32, n: *mut u32,
                       consumed: *mut u64, max_rec_bytes: *mut u32) -> c_int;
    // run.rs batch map
    pub fn bk_device_count() -> c_int;
    pub fn bk_warmup(device: c_int) -> c_int;
    pub fn bk_ctx_create(device: c_int, p1: *const BkProgram, p2: *const BkProgram, flags: u32, status: *mut c_int,
                         err: *mut c_char, errcap: usize) -> *mut BkCtx;
    pub fn bk_ctx_set_host_threads(ctx: *mut BkCtx, n: c_int) -> c_int;
    pub fn bk_ctx_destroy(ctx: *mut BkCtx);
    pub fn bk_last_error(ctx: *const BkCtx) -> *const c_char;
    pub fn bk_host_alloc(bytes: usize) -> *mut u8;
    pub fn bk_host_free(p: *mut u8);
    pub fn bk_out_bound(ctx: *const BkCtx, mate: c_int, n: u32, text_len: u64) -> u64;
    pub fn bk_ctx_nslots(ctx: *const BkCtx) -> c_int;
    pub fn bk_submit(ctx: *mut BkCtx, slot: c_int, n: u32, in1: *const BkMateIn, in2: *const BkMateIn) -> c_int;
    pub fn bk_submit_text(ctx: *mut BkCtx, slot: c_int, n: u32, text1: *const u8, len1: u64, text2: *const u8, len2: u64,
                          max_rec_hint: u32) -> c_int;
    pub fn bk_wait(ctx: *mut BkCtx, slot: c_int, out1: *mut u8, cap1: u64, out2: *mut u8, cap2: u64,
                   res: *mut BkResult) -> c_int;
    pub fn bk_extract_host(ctx: *mut BkCtx, n: u32, in1: *const BkMateIn, in2: *const BkMateIn, out1: *mut u8,
                           cap1: u64, out2: *mut u8, cap2: u64, res: *mut BkResult) -> c_int;
    // parse.rs
    pub fn bk_reverse_complement(ctx: *mut BkCtx, seq: *const u8, n: usize, out: *mut u8) -> c_int;
    // run::run
    pub fn bk_run(args: *const BkRunArgs, err: *mut c_char, errcap: usize) -> c_int;
}
