// build.rs -- builds libgsv_cuda.so with nvcc for sm_100a (the engine's own Makefile) and links it.
//
//   GSV_CUDA_LIB_DIR  directory that already holds libgsv_cuda.so (skips the build)
//   NVCC              nvcc to use (default: /usr/local/cuda/bin/nvcc, forwarded to the Makefile)
use std::{env, path::PathBuf, process::Command};

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let pkg = manifest.join("..").join("garbled-snark-verifier_b200");
    let lib_dir = match env::var("GSV_CUDA_LIB_DIR") {
        Ok(d) => PathBuf::from(d),
        Err(_) => {
            // nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... (csrc/Makefile)
            let status = Command::new("make").arg("-C").arg(pkg.join("csrc")).status().expect("make not found");
            assert!(status.success(), "building libgsv_cuda.so failed");
            pkg.clone()
        }
    };
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=gsv_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib_dir.display());
    println!("cargo:rerun-if-changed={}", manifest.join("../include/gsv_cuda.h").display());
    println!("cargo:rerun-if-changed={}", pkg.join("csrc").display());
    println!("cargo:rerun-if-env-changed=GSV_CUDA_LIB_DIR");
}
