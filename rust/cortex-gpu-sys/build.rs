// Compiles the CUDA sources of cortex_b200 for sm_100a with nvcc and links the result.
// Mirrors cortex_b200/build.py (same flags); there is no other backend and no CPU path.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("cortex_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libcortex_gpu.so");
    let mut cmd = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()));
    cmd.args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "--expt-relaxed-constexpr"])
        .arg("-I").arg(root.join("include")).arg("-I").arg(&csrc)
        .arg("-o").arg(&lib);
    for f in ["cx_index.cu", "cx_search.cu", "cx_exact.cu", "cx_stream.cu", "cx_tensor.cu", "cx_select.cu", "cx_merge.cu"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
        cmd.arg(csrc.join(f));
    }
    let st = cmd.status().expect("nvcc not found: cortex-gpu has no CPU fallback");
    assert!(st.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=cortex_gpu");
}
