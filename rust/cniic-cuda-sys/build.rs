// build.rs -- compiles the .cu files with nvcc for sm_100a and links the result (north_star: "a build.rs compiles the
// .cu files with nvcc for sm_100a and links them").  NOT built in this repository (no cargo in the image).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../cniic_b200/csrc");
    let mut objs = Vec::new();
    for name in ["api", "kmeans", "sort", "stages", "codec", "synth"] {
        let obj = out.join(format!("{name}.o"));
        let status = Command::new("nvcc")
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
                   "--expt-relaxed-constexpr", "-c", "-o"])
            .arg(&obj)
            .arg(csrc.join(format!("{name}.cu")))
            .status()
            .expect("nvcc not found");
        assert!(status.success(), "nvcc failed on {name}.cu");
        objs.push(obj);
    }
    let lib = out.join("libcniic_b200.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=cniic_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
    println!("cargo:rerun-if-changed=../../cniic_b200/csrc");
    println!("cargo:rerun-if-changed=../../include/cniic_b200.h");
}
