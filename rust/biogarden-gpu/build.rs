// Builds the CUDA library with nvcc for sm_100a (same command as biogarden_b200/csrc/Makefile) and links it.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("biogarden_b200/csrc");
    let status = Command::new("make").arg("-C").arg(&csrc).status().expect("make (nvcc) failed to start");
    assert!(status.success(), "nvcc build of libbgalign.so failed");
    println!("cargo:rustc-link-search=native={}", root.join("biogarden_b200").display());
    println!("cargo:rustc-link-lib=dylib=bgalign");
    println!("cargo:rerun-if-changed={}", csrc.display());
}
