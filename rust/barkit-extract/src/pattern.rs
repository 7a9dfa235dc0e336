//! pattern.rs of the reference on top of `bk_compile` / `bk_expand`.
use std::ffi::{CStr, CString};
use std::fmt;

use crate::error::Error;
use crate::ffi;

pub struct BarcodePattern {
    adapter_pattern: String,
    max_error: usize,
}

fn call_text(f: impl Fn(*mut std::os::raw::c_char, usize) -> i32) -> Result<String, Error> {
    let mut cap = 1usize << 16;
    loop {
        let mut buf = vec![0u8; cap];
        let rc = f(buf.as_mut_ptr() as *mut _, cap);
        if rc == ffi::BK_E_CAPACITY && cap < (1 << 28) {
            cap *= 16;
            continue;
        }
        let text = unsafe { CStr::from_ptr(buf.as_ptr() as *const _) }.to_string_lossy().into_owned();
        return if rc == ffi::BK_OK { Ok(text) } else { Err(Error::from_status(rc, text)) };
    }
}

impl BarcodePattern {
    pub fn new(pattern: &str, max_error: &usize) -> Result<Self, Error> {
        Ok(BarcodePattern { adapter_pattern: pattern.to_owned(), max_error: *max_error })
    }

    /// pattern.rs:40-79
    pub fn get_sequence_with_errors(&self, sequence: &str) -> Result<Vec<String>, Error> {
        let seq = CString::new(sequence).map_err(|e| Error::Regex(e.to_string()))?;
        let mut n = 0usize;
        let text = call_text(|out, cap| unsafe {
            ffi::bk_sequence_with_errors(seq.as_ptr(), self.max_error, out, cap, &mut n)
        })?;
        Ok(if n == 0 { vec![] } else { text.split('\n').map(str::to_owned).collect() })
    }

    /// pattern.rs:93-109
    pub fn get_pattern_with_errors(&self) -> Result<String, Error> {
        let pat = CString::new(self.adapter_pattern.as_str()).map_err(|e| Error::Regex(e.to_string()))?;
        call_text(|out, cap| unsafe { ffi::bk_expand(pat.as_ptr(), self.max_error, out, cap) })
    }
}

#[derive(Clone, Debug, PartialEq, Eq)]
pub enum BarcodeType {
    Umi,
    Sample,
    Cell,
}

impl BarcodeType {
    /// pattern.rs:137-144
    pub fn parse_type(name: &str) -> Result<Self, Error> {
        match name {
            "UMI" => Ok(BarcodeType::Umi),
            "SB" => Ok(BarcodeType::Sample),
            "CB" => Ok(BarcodeType::Cell),
            _ => Err(Error::UnexpectedCaptureGroupName(name.to_owned())),
        }
    }
}

impl fmt::Display for BarcodeType {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        f.write_str(match self {
            BarcodeType::Umi => "UMI",
            BarcodeType::Sample => "SB",
            BarcodeType::Cell => "CB",
        })
    }
}

/// pattern.rs:161-235: the compiled pattern.  Matching itself happens on the device, a batch at a time
/// (`parse::BatchExtractor`); there is no per-read `get_captures` on the host.
pub struct BarcodeRegex {
    pub(crate) program: *mut ffi::BkProgram,
    barcode_types: Vec<BarcodeType>,
}

unsafe impl Send for BarcodeRegex {}
unsafe impl Sync for BarcodeRegex {} // bk_program is immutable after bk_compile

impl BarcodeRegex {
    pub fn new(pattern: &str, max_error: usize) -> Result<Self, Error> {
        let pat = CString::new(pattern).map_err(|e| Error::Regex(e.to_string()))?;
        let mut status = 0;
        let mut err = vec![0u8; 2048];
        let program =
            unsafe { ffi::bk_compile(pat.as_ptr(), max_error, &mut status, err.as_mut_ptr() as *mut _, err.len()) };
        if program.is_null() {
            let msg = unsafe { CStr::from_ptr(err.as_ptr() as *const _) }.to_string_lossy().into_owned();
            return Err(Error::from_status(status, msg));
        }
        let n = unsafe { ffi::bk_program_ncaps(program) };
        let mut barcode_types = Vec::with_capacity(n as usize);
        for i in 0..n {
            let name = unsafe { CStr::from_ptr(ffi::bk_program_cap_name(program, i)) }.to_string_lossy();
            barcode_types.push(BarcodeType::parse_type(&name)?);
        }
        Ok(BarcodeRegex { program, barcode_types })
    }

    /// pattern.rs:232-234
    pub fn get_barcode_types(&self) -> Vec<BarcodeType> {
        self.barcode_types.clone()
    }
}

impl Drop for BarcodeRegex {
    fn drop(&mut self) {
        unsafe { ffi::bk_program_destroy(self.program) }
    }
}
