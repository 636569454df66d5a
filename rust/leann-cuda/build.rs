// build.rs — compiles the CUDA sources of this repository with nvcc for sm_100a and links them.
// SOURCE ONLY: cargo/rustc are not available in the build image of this repository, so this crate is
// not compiled there; the C ABI it binds (include/leann_cuda.h) is exercised from Python/ctypes instead.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("leann_rs_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libleann_cuda.so");
    let mut sources = Vec::new();
    for e in std::fs::read_dir(&csrc).unwrap() {
        let p = e.unwrap().path();
        match p.extension().and_then(|s| s.to_str()) {
            Some("cu") | Some("cpp") => sources.push(p),
            _ => {}
        }
        println!("cargo:rerun-if-changed={}", csrc.display());
    }
    let status = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()))
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo"])
        .args(["-Xcompiler", "-fPIC,-pthread", "--expt-relaxed-constexpr", "-x", "cu", "-shared", "-o"])
        .arg(&lib)
        .args(&sources)
        .status()
        .expect("nvcc not found: the leann-cuda backend has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=leann_cuda");
}
