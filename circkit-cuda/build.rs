// Builds libcirckit_b200.so with nvcc for sm_100a (nvcc cross-compiles without a GPU) and the host packer with the C++
// compiler, exactly as circkit_b200/build.py does, and links it.  CIRCKIT_B200_CSRC overrides where the sources are
// (default: ../circkit_b200/csrc next to this crate); CIRCKIT_B200_LIB_DIR skips the build and links a prebuilt library.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    println!("cargo:rerun-if-env-changed=CIRCKIT_B200_LIB_DIR");
    println!("cargo:rerun-if-env-changed=CIRCKIT_B200_CSRC");
    if let Ok(dir) = env::var("CIRCKIT_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=circkit_b200");
        return;
    }
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = env::var("CIRCKIT_B200_CSRC").map(PathBuf::from).unwrap_or_else(|_| manifest.join("../circkit_b200/csrc"));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", manifest.join("../include/circkit_b200.h").display());

    // the host packer (ck_pack2_host): plain C++ with run-time-dispatched AVX2 / BMI2 paths
    let obj = out.join("ck_host_pack.o");
    let cxx = env::var("CXX").unwrap_or_else(|_| "g++".into());
    let st = Command::new(&cxx)
        .args(["-O3", "-std=c++17", "-fPIC", "-pthread", "-c", "-o"])
        .arg(&obj)
        .arg(csrc.join("ck_host_pack.cpp"))
        .status()
        .expect("C++ compiler for ck_host_pack.cpp");
    assert!(st.success(), "ck_host_pack.cpp failed to compile");

    // every kernel + the C ABI: one translation unit, sm_100a only
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let st = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-o"])
        .arg(out.join("libcirckit_b200.so"))
        .arg(csrc.join("ck_lib.cu"))
        .arg(&obj)
        .args(["-Xcompiler", "-pthread"])
        .status()
        .expect("nvcc (CUDA 12.8+ for sm_100a)");
    assert!(st.success(), "ck_lib.cu failed to compile");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=circkit_b200");
}
