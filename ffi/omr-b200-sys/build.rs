// Links libomr_b200.so.  OMR_B200_LIB_DIR = directory that holds it (default: ../../tfhe-omr_b200/lib relative to this crate,
// i.e. the in-tree build of `python -c "import __graft_entry__ as g; g.build()"`).
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("OMR_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../tfhe-omr_b200/lib")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=omr_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=OMR_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/omr_b200.h");
}
