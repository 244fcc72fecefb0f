//! # omr-b200-sys
//!
//! Rust binding of `libomr_b200.so`, the B200 (sm_100a) implementation of InstantOMR's detection hot path, for the reference
//! workspace `xiangxiecrypto/tfhe-omr`.
//!
//! * [`sys`] — the raw C ABI: every entry point, struct and constant of `include/omr_b200.h` (generated from the header).
//! * [`blob`] — the versioned flat containers (`*.omrb`) that the library, the CPU oracle and the Python host side exchange.
//! * [`GpuDetector`] (feature `reference`) — a drop-in for `omr_core::Detector` with the same method signatures
//!   (`omr_core/src/detector.rs:85, 135-138, 169-175, 223-227, 341-351`), plus the batched `detect_batch` that replaces
//!   `clues_list.par_iter().map(|c| detector.detect(c))` (`omr_core/examples/omr.rs:160-164`).
//!
//! Interop conventions: keys go in and ciphertexts come out in **coefficient form** (`OMR_KEYS_COEFF`,
//! `OMR_OUT_COEFF`): every polynomial is passed through Primus-fhe's own NTT table at the boundary, so nothing depends on the
//! two libraries agreeing on the root of unity or on the ordering of NTT-domain vectors.  There is no CPU fallback: without a
//! CUDA device `GpuDetector::new` panics with the library's message.
//!
//! This crate could not be compiled where it was written (no Rust toolchain); `src/sys.rs` is generated from the header and
//! checked against it by `tests/test_ffi_crate.py`.  Everything that touches a Primus-fhe accessor whose name could not be
//! verified lives in `src/flatten.rs` and is marked `[UPSTREAM]`.

pub mod blob;
pub mod sys;

#[cfg(feature = "reference")]
pub mod flatten;
#[cfg(feature = "reference")]
mod detector;
#[cfg(feature = "reference")]
pub use detector::{DetectTimeInfoPerMessage, GpuDetector};

use std::ffi::CStr;

/// Error of a C-ABI call: status code (`OMR_ERR_*`) and the library's message.
#[derive(Debug, Clone)]
pub struct OmrGpuError {
    pub status: i32,
    pub message: String,
}

impl std::fmt::Display for OmrGpuError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "omr_b200 status {}: {}", self.status, self.message)
    }
}
impl std::error::Error for OmrGpuError {}

/// `omr_last_error(ctx)` (NULL = the last context-free failure) as a `String`.
pub fn last_error(ctx: *const sys::OmrCtx) -> String {
    // SAFETY: omr_last_error returns a NUL-terminated string owned by the library (never NULL)
    unsafe {
        let p = sys::omr_last_error(ctx);
        if p.is_null() {
            String::new()
        } else {
            CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    }
}

pub(crate) fn check(status: i32, ctx: *const sys::OmrCtx) -> Result<(), OmrGpuError> {
    if status == sys::OMR_OK {
        Ok(())
    } else {
        Err(OmrGpuError { status, message: last_error(ctx) })
    }
}
