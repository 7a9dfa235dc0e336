//! Same module paths as the reference crate (barkit-extract/src/lib.rs:1-6); `ffi` is the only addition.
pub mod error;
pub mod fastq;
pub mod ffi;
pub mod logger;
pub mod parse;
pub mod pattern;
pub mod run;
