// bioshell-seq/build.rs -- addition for the B200 aligner (SOURCE ONLY: there is no Rust
// toolchain in the build image, so this file has not been compiled; the tested boundary is
// the C ABI in include/bioshell_align.h).
//
// The reference's build.rs only exports BUILD_TIME / GIT_COMMIT_MD5 (bioshell-seq/build.rs:1-29);
// this adds the nvcc step the north star asks for and links the resulting static library.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from("cuda"); // bioshell_b200/csrc copied next to Cargo.toml
    let obj = out.join("bsa_api.o");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let ok = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-Xcompiler", "-fPIC", "-c", "-o"])
        .arg(&obj)
        .arg(csrc.join("bsa_api.cu"))
        .status()
        .expect("nvcc not found")
        .success();
    assert!(ok, "nvcc failed");
    let lib = out.join("libbioshell_align.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).arg(&obj).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=bioshell_align");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rerun-if-changed=cuda/bsa_api.cu");
    println!("cargo:rerun-if-changed=cuda/gotoh_kernels.cuh");
}
