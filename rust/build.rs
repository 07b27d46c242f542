// build.rs — replacement for the reference's build script (build.rs:1-118): compiles the B200-native library for
// sm_100a ONLY (no other arch, no PTX fallback, no arch probing) and links it as the static library `ntt_cuda`
// that src/ntt.rs:95 expects.  NOT COMPILED HERE (no Rust toolchain in this image); toyni_b200/build.py performs the
// same nvcc + ar steps and tests/test_abi_symbols.py checks the result.
use std::env;
use std::path::PathBuf;
use std::process::Command;

// Every .cu file under cuda/ (= toyni_b200/csrc of this repository) is one translation unit of the library.  The list
// is read from the directory, not written down twice: toyni_b200/build.py keeps an explicit list and
// tests/test_abi_symbols.py::test_build_lists_agree checks that it names exactly the .cu files that exist.
fn sources(dir: &PathBuf) -> Vec<String> {
    let mut v: Vec<String> = std::fs::read_dir(dir)
        .expect("cuda/ directory missing")
        .filter_map(|e| e.ok())
        .map(|e| e.file_name().to_string_lossy().into_owned())
        .filter(|n| n.ends_with(".cu"))
        .collect();
    v.sort();
    assert!(!v.is_empty(), "no .cu sources under cuda/");
    v
}

fn main() {
    println!("cargo:rerun-if-changed=cuda");
    if env::var("CARGO_FEATURE_CUDA").is_err() {
        return;
    }
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string());
    // Unlike the reference (build.rs:34-37), a missing nvcc is a hard error: src/ntt.rs links `ntt_cuda` whenever
    // the feature is on, so returning early would only move the failure to link time.
    let ok = Command::new(&nvcc).arg("--version").output().map(|o| o.status.success()).unwrap_or(false);
    assert!(ok, "feature `cuda` needs nvcc (CUDA >= 12.8 for sm_100a)");

    let out_dir = PathBuf::from(env::var("OUT_DIR").unwrap());
    let src_dir = PathBuf::from("cuda"); // toyni_b200/csrc/*, toyni_b200/host/*.hpp and include/*.h of this repository, flat
    let mut objects = Vec::new();
    for src in sources(&src_dir).iter() {
        let obj = out_dir.join(src.replace(".cu", ".o"));
        let status = Command::new(&nvcc)
            .args(["-c", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"])
            .args(["-gencode", "arch=compute_100a,code=sm_100a"]) // sm_100a SASS only
            .arg(src_dir.join(src))
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("failed to run nvcc");
        assert!(status.success(), "nvcc failed on {src}");
        objects.push(obj);
    }
    let lib = out_dir.join("libntt_cuda.a");
    let _ = std::fs::remove_file(&lib);
    let status = Command::new("ar").arg("rcs").arg(&lib).args(&objects).status().expect("failed to run ar");
    assert!(status.success(), "ar failed");

    println!("cargo:rustc-link-search=native={}", out_dir.display());
    println!("cargo:rustc-link-lib=static=ntt_cuda");
    let cuda_home = env::var("CUDA_HOME").or_else(|_| env::var("CUDA_PATH")).unwrap_or_else(|_| "/usr/local/cuda".into());
    println!("cargo:rustc-link-search=native={cuda_home}/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    // The single-process multi-GPU entry points (bb_mg_*, header section 4) use CUDA peer access only; nothing else to
    // link.  A host that runs one process per GPU instead links NCCL itself for its exchange:
    // println!("cargo:rustc-link-lib=dylib=nccl");
}
