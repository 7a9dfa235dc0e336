//! run.rs of the reference: the same `run` signature (run.rs:10-25), forwarded to `bk_run`, which holds the
//! two loops (run.rs:83-146, :198-276) on the device side of the C ABI.
use std::ffi::{CStr, CString};

use crate::fastq::CompressionType;
use crate::ffi;

#[allow(clippy::too_many_arguments)]
pub fn run(
    fq1: String,
    fq2: Option<String>,
    pattern1: Option<String>,
    pattern2: Option<String>,
    out_fq1: String,
    out_fq2: Option<String>,
    max_memory: Option<usize>,
    threads: usize,
    rc_barcodes: bool,
    skip_trimming: bool,
    max_error: usize,
    output_compression: CompressionType,
    quiet: bool,
    force: bool,
) {
    let c = |s: &String| CString::new(s.as_str()).expect("argument contains a NUL byte");
    let fq1 = c(&fq1);
    let out_fq1 = c(&out_fq1);
    let (fq2, pattern1, pattern2, out_fq2) =
        (fq2.as_ref().map(c), pattern1.as_ref().map(c), pattern2.as_ref().map(c), out_fq2.as_ref().map(c));
    let p = |o: &Option<CString>| o.as_ref().map_or(std::ptr::null(), |s| s.as_ptr());
    let args = ffi::BkRunArgs {
        fq1: fq1.as_ptr(),
        fq2: p(&fq2),
        pattern1: p(&pattern1),
        pattern2: p(&pattern2),
        out_fq1: out_fq1.as_ptr(),
        out_fq2: p(&out_fq2),
        max_memory_mb: max_memory.unwrap_or(0),
        threads,
        rc_barcodes: rc_barcodes as i32,
        skip_trimming: skip_trimming as i32,
        max_error,
        compression: output_compression.code(),
        quiet: quiet as i32,
        force: force as i32,
        n_gpus: 0,
        verbose: 0,
        ctx_flags: 0,
    };
    let mut err = vec![0u8; 2048];
    let rc = unsafe { ffi::bk_run(&args, err.as_mut_ptr() as *mut _, err.len()) };
    if rc != ffi::BK_OK {
        // run.rs:102-118 / :220-245: eprintln!("{}", e); std::process::exit(1)
        eprintln!("{}", unsafe { CStr::from_ptr(err.as_ptr() as *const _) }.to_string_lossy());
        std::process::exit(1);
    }
}
