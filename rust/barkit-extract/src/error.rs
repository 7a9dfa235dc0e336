//! error.rs of the reference, minus the variants that wrapped its regex engines' own error types: the
//! library reports those as text with the same wording (`Regex error: ...`).
use thiserror::Error;

#[derive(Error, Debug, Clone)]
pub enum Error {
    #[error("{0}")]
    Regex(String), // "Regex error: ..." / "Fancy regex error: ..." as produced by bk_compile
    #[error("{0} capture group not found in your pattern")]
    BarcodeCaptureGroupNotFound(String),
    #[error("Provided unexpected barcode capture group {0}")]
    UnexpectedCaptureGroupName(String),
    #[error("I/O error: {0}")]
    IO(String),
    #[error("No match")]
    PatternNotMatched,
    #[error("Failed to choose permutation mask")]
    PermutationMaskSize,
    /// valid for the reference's regex engine, outside what the B200 build lowers (status -2)
    #[error("{0}")]
    Unsupported(String),
    /// CUDA / capacity / argument errors of the library (status -3, -4, -6, -7)
    #[error("{0}")]
    Device(String),
}

impl Error {
    /// status + message of the C ABI -> the reference's error classes (error.rs:5-22)
    pub(crate) fn from_status(status: i32, msg: String) -> Self {
        use crate::ffi::*;
        match status {
            BK_E_PATTERN => {
                if let Some(name) = msg.strip_prefix("Provided unexpected barcode capture group ") {
                    Error::UnexpectedCaptureGroupName(name.to_string())
                } else if let Some(p) = msg.strip_suffix(" capture group not found in your pattern") {
                    Error::BarcodeCaptureGroupNotFound(p.to_string())
                } else if msg == "Failed to choose permutation mask" {
                    Error::PermutationMaskSize
                } else {
                    Error::Regex(msg)
                }
            }
            BK_E_UNSUPPORTED => Error::Unsupported(msg),
            BK_E_IO => Error::IO(msg.strip_prefix("I/O error: ").unwrap_or(&msg).to_string()),
            _ => Error::Device(msg),
        }
    }
}
