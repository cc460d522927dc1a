// Builds libkgr_msm.so for sm_100a with the repo's own Makefile (nvcc -gencode arch=compute_100a,code=sm_100a)
// and links it.  Untested: no Rust toolchain exists in the image this repository is developed in.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("kogarashi_b200/csrc");
    let status = Command::new("make").arg("-C").arg(&csrc).status().expect("make not found");
    assert!(status.success(), "building libkgr_msm.so failed");
    println!("cargo:rustc-link-search=native={}", root.join("kogarashi_b200").display());
    println!("cargo:rustc-link-lib=dylib=kgr_msm");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include/kgr_msm.h").display());
}
