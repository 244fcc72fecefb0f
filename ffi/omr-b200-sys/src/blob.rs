//! Safe wrappers of `omr_blob_*`: the `*.omrb` containers of `include/omr_b200.h`.
//! 64-byte header (`"OMRB200\0"`, version, kind, count, index0, aux, payload bytes, domain) + the raw arrays of the kind.

use std::{ffi::CString, os::raw::c_void, path::Path};

use crate::{check, sys, OmrGpuError};

pub use sys::{
    OMR_BLOB_CLUES as CLUES, OMR_BLOB_CLUE_KEY as CLUE_KEY, OMR_BLOB_DETECTION_KEY as DETECTION_KEY, OMR_BLOB_DIGEST as DIGEST,
    OMR_BLOB_LWE2 as LWE2, OMR_BLOB_PAYLOADS as PAYLOADS, OMR_BLOB_PERTINENCY_VECTOR as PERTINENCY_VECTOR, OMR_BLOB_RLWE1 as RLWE1,
    OMR_BLOB_RLWE2 as RLWE2, OMR_BLOB_SECRET_KEY as SECRET_KEY, OMR_OUT_COEFF as DOMAIN_COEFF, OMR_OUT_NTT_NATIVE as DOMAIN_NTT_NATIVE,
};

fn cpath(path: &Path) -> CString {
    CString::new(path.to_string_lossy().as_bytes()).expect("path contains a NUL byte")
}

/// One array of a blob, as bytes in memory order (little-endian hosts only, like the library).
pub fn bytes_of<T: Copy>(v: &[T]) -> &[u8] {
    // SAFETY: plain-old-data slices reinterpret as bytes
    unsafe { std::slice::from_raw_parts(v.as_ptr() as *const u8, std::mem::size_of_val(v)) }
}

/// `omr_blob_write`: `arrays` in the declaration order of the kind; their sizes are checked against `count`.
pub fn write(path: &Path, kind: u32, count: u64, index0: u64, aux: u64, domain: u32, arrays: &[&[u8]]) -> Result<(), OmrGpuError> {
    for (i, a) in arrays.iter().enumerate() {
        // SAFETY: pure function of its arguments
        let want = unsafe { sys::omr_blob_field_bytes(kind, i as u32, count) };
        assert_eq!(a.len(), want, "blob kind {kind}: array {i} has {} bytes, expected {want}", a.len());
    }
    let ptrs: Vec<*const c_void> = arrays.iter().map(|a| a.as_ptr() as *const c_void).collect();
    let p = cpath(path);
    // SAFETY: pointers stay valid for the call; sizes were checked above
    let st = unsafe { sys::omr_blob_write(p.as_ptr(), kind, count, index0, aux, domain, ptrs.as_ptr(), ptrs.len() as u32) };
    check(st, std::ptr::null())
}

/// `omr_blob_read_header` + `omr_blob_read`: the header and the arrays as byte vectors.
pub fn read(path: &Path) -> Result<(sys::OmrBlobHeader, Vec<Vec<u8>>), OmrGpuError> {
    let p = cpath(path);
    let mut hdr = sys::OmrBlobHeader { version: 0, kind: 0, count: 0, index0: 0, aux: 0, payload_bytes: 0, domain: 0, reserved: [0; 3] };
    // SAFETY: hdr is a valid out-pointer
    check(unsafe { sys::omr_blob_read_header(p.as_ptr(), &mut hdr) }, std::ptr::null())?;
    // SAFETY: pure functions
    let n = unsafe { sys::omr_blob_field_count(hdr.kind) };
    let mut bufs: Vec<Vec<u8>> = (0..n).map(|i| vec![0u8; unsafe { sys::omr_blob_field_bytes(hdr.kind, i, hdr.count) }]).collect();
    let ptrs: Vec<*mut c_void> = bufs.iter_mut().map(|b| b.as_mut_ptr() as *mut c_void).collect();
    // SAFETY: each buffer has exactly the size the library will write
    check(unsafe { sys::omr_blob_read(p.as_ptr(), &mut hdr, ptrs.as_ptr(), n) }, std::ptr::null())?;
    Ok((hdr, bufs))
}
