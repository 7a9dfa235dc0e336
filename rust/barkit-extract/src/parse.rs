//! parse.rs of the reference, batch-shaped: `BarcodeParser::parse_barcodes` (parse.rs:58-68) worked on one
//! record; the device works on a RecordSet's whole buffer, so the unit here is the batch map
//! `parse_se_reads` / `parse_pe_reads` (run.rs:63-80, :163-195) followed by the writer's framing
//! (fastq.rs:259-263): FASTQ text in, FASTQ text out, order preserved.
use std::ffi::CStr;

use crate::error::Error;
use crate::ffi;
use crate::pattern::BarcodeRegex;

pub struct BatchExtractor {
    ctx: *mut ffi::BkCtx,
    paired: bool,
}

pub struct BatchOutput {
    pub out1: Vec<u8>,
    pub out2: Vec<u8>,
    pub stats: ffi::BkResult,
}

fn tokenize(text: &[u8]) -> Result<(Vec<u32>, u32), Error> {
    let cap = (text.iter().filter(|&&b| b == b'\n').count() / 4 + 2) as u32;
    let mut off = vec![0u32; cap as usize + 1];
    let (mut n, mut consumed, mut maxrec) = (0u32, 0u64, 0u32);
    let rc = unsafe {
        ffi::bk_tokenize(text.as_ptr(), text.len() as u64, 1, off.as_mut_ptr(), cap, &mut n, &mut consumed, &mut maxrec)
    };
    if rc != ffi::BK_OK || consumed != text.len() as u64 {
        return Err(Error::Device("malformed FASTQ text".into()));
    }
    off.truncate(n as usize + 1);
    Ok((off, maxrec.max(1)))
}

impl BatchExtractor {
    /// `BarcodeParser::new` (parse.rs:46-56) for a whole run: either pattern may be absent in paired mode.
    pub fn new(device: i32, barcode1: Option<&BarcodeRegex>, barcode2: Option<&BarcodeRegex>, paired: bool,
               skip_trimming: bool, rc_barcodes: bool) -> Result<Self, Error> {
        let flags = if paired { ffi::BK_FLAG_PAIRED } else { 0 }
            | if skip_trimming { ffi::BK_FLAG_SKIP_TRIM } else { 0 }
            | if rc_barcodes { ffi::BK_FLAG_RC } else { 0 };
        let mut status = 0;
        let mut err = vec![0u8; 1024];
        let ctx = unsafe {
            ffi::bk_ctx_create(device, barcode1.map_or(std::ptr::null(), |b| b.program as *const _),
                               barcode2.map_or(std::ptr::null(), |b| b.program as *const _), flags, &mut status,
                               err.as_mut_ptr() as *mut _, err.len())
        };
        if ctx.is_null() {
            let msg = unsafe { CStr::from_ptr(err.as_ptr() as *const _) }.to_string_lossy().into_owned();
            return Err(Error::from_status(status, msg));
        }
        Ok(BatchExtractor { ctx, paired })
    }

    /// One batch: the text of a RecordSet per mate (`fq2` None in single-end mode).  Pairs are formed by record
    /// index; the shorter side bounds the batch (run.rs:172-173).
    pub fn extract(&mut self, fq1: &[u8], fq2: Option<&[u8]>) -> Result<BatchOutput, Error> {
        assert_eq!(self.paired, fq2.is_some());
        let pad = |t: &[u8]| { let mut v = t.to_vec(); v.resize(t.len() + 64, 0); v }; // 16 B of slack behind the text
        let (t1, t2) = (pad(fq1), fq2.map(pad));
        let (off1, max1) = tokenize(fq1)?;
        let tok2 = match fq2 { Some(t) => Some(tokenize(t)?), None => None };
        let n = tok2.as_ref().map_or(off1.len() - 1, |(o, _)| (o.len() - 1).min(off1.len() - 1));
        let in1 = ffi::BkMateIn { text: t1.as_ptr(), rec_off: off1.as_ptr(), text_len: off1[n] as u64,
                                  max_rec_bytes: max1, reserved: 0 };
        let in2 = tok2.as_ref().map(|(o, m)| ffi::BkMateIn {
            text: t2.as_ref().unwrap().as_ptr(), rec_off: o.as_ptr(), text_len: o[n] as u64, max_rec_bytes: *m,
            reserved: 0 });
        let cap1 = unsafe { ffi::bk_out_bound(self.ctx, 1, n as u32, in1.text_len) };
        let cap2 = in2.as_ref().map_or(0, |m| unsafe { ffi::bk_out_bound(self.ctx, 2, n as u32, m.text_len) });
        let (mut out1, mut out2) = (vec![0u8; cap1 as usize + 16], vec![0u8; cap2 as usize + 16]);
        let mut stats = ffi::BkResult::default();
        let rc = unsafe {
            ffi::bk_extract_host(self.ctx, n as u32, &in1, in2.as_ref().map_or(std::ptr::null(), |m| m as *const _),
                                 out1.as_mut_ptr(), out1.len() as u64,
                                 if self.paired { out2.as_mut_ptr() } else { std::ptr::null_mut() }, out2.len() as u64,
                                 &mut stats)
        };
        if rc != ffi::BK_OK {
            let msg = unsafe { CStr::from_ptr(ffi::bk_last_error(self.ctx)) }.to_string_lossy().into_owned();
            return Err(Error::from_status(rc, msg));
        }
        out1.truncate(stats.out1_len as usize);
        out2.truncate(stats.out2_len as usize);
        Ok(BatchOutput { out1, out2, stats })
    }

    /// parse.rs:173-179, computed with the kernel's own complement table
    pub fn get_reverse_complement(&mut self, sequence: &[u8]) -> Vec<u8> {
        let mut out = vec![0u8; sequence.len()];
        unsafe { ffi::bk_reverse_complement(self.ctx, sequence.as_ptr(), sequence.len(), out.as_mut_ptr()) };
        out
    }
}

impl Drop for BatchExtractor {
    fn drop(&mut self) {
        unsafe { ffi::bk_ctx_destroy(self.ctx) }
    }
}
