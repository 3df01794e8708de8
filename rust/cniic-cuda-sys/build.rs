// build.rs -- compiles EVERY cniic_b200/csrc/*.cu with nvcc for sm_100a and links the result (north_star: "a build.rs compiles
// the .cu files with nvcc for sm_100a and links them").  The file list is read from the directory, exactly like
// cniic_b200/csrc/Makefile's $(wildcard *.cu), so a new translation unit can never be forgotten here
// (tests/test_cpu_host.py checks that no list is hard-coded).  NOT built in this repository (no cargo in the image).
use std::{env, fs, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../cniic_b200/csrc");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let mut sources: Vec<PathBuf> = fs::read_dir(&csrc)
        .expect("cniic_b200/csrc not found")
        .filter_map(|e| e.ok().map(|e| e.path()))
        .filter(|p| p.extension().map_or(false, |x| x == "cu"))
        .collect();
    sources.sort();
    assert!(!sources.is_empty(), "no .cu files under {}", csrc.display());
    let mut objs = Vec::new();
    for src in &sources {
        let stem = src.file_stem().unwrap().to_string_lossy().into_owned();
        let obj = out.join(format!("{stem}.o"));
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
                   "--expt-relaxed-constexpr", "-c", "-o"])
            .arg(&obj)
            .arg(src)
            .status()
            .expect("nvcc not found (set NVCC)");
        assert!(status.success(), "nvcc failed on {}", src.display());
        objs.push(obj);
    }
    let lib = out.join("libcniic_b200.a");
    let _ = fs::remove_file(&lib);
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=cniic_b200");
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");   // NCCL is dlopen()ed at run time (multi-GPU contexts only)
    println!("cargo:rerun-if-changed=../../cniic_b200/csrc");
    println!("cargo:rerun-if-changed=../../include/cniic_b200.h");
    println!("cargo:rerun-if-env-changed=NVCC");
}
