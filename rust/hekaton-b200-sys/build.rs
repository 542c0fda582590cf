//! Locates `libhekaton_b200.so`.  The library is built by this repository (`python __graft_entry__.py`, or
//! `make -C hekaton_system_b200/csrc`) into `hekaton_system_b200/lib/`; point `HEKATON_B200_LIB_DIR` at that
//! directory (or at wherever the .so was installed).  CUDA's runtime is linked into the .so itself, so nothing
//! else is needed at link time; at run time the directory must be on the loader path (an rpath is emitted).
use std::{env, path::PathBuf};

fn main() {
    println!("cargo:rerun-if-env-changed=HEKATON_B200_LIB_DIR");
    let dir = env::var_os("HEKATON_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|| {
        // default: this crate sits in <repo>/rust/hekaton-b200-sys
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../hekaton_system_b200/lib")
    });
    let dir = dir.canonicalize().unwrap_or(dir);
    if !dir.join("libhekaton_b200.so").exists() {
        panic!(
            "libhekaton_b200.so not found in {} -- build it first (make -C hekaton_system_b200/csrc) or set \
             HEKATON_B200_LIB_DIR; there is no CPU fallback",
            dir.display()
        );
    }
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=hekaton_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:lib_dir={}", dir.display());
}
