// Replacement for the reference's 3-line build.rs (build.rs:1-3): keeps shadow-rs and, with
// `--features cuda`, compiles the B200 backend with nvcc and links it.
//
// UNBUILT ARTEFACT (no Rust toolchain in the development image).  The nvcc command line is the one
// `mpc-iris-code_b200/build.py` runs; `B200_SRC` points at this repository's `mpc-iris-code_b200/csrc`.
fn main() -> shadow_rs::SdResult<()> {
    if std::env::var("CARGO_FEATURE_CUDA").is_ok() {
        let out = std::env::var("OUT_DIR").unwrap();
        let src = std::env::var("B200_SRC").unwrap_or_else(|_| "b200/csrc".into());
        let lib = format!("{out}/libiris_b200.so");
        let status = std::process::Command::new("nvcc")
            .args([
                "-O3",
                "-std=c++17",
                "-gencode",
                "arch=compute_100a,code=sm_100a",
                "-lineinfo",
                "-Xcompiler",
                "-fPIC",
                "--expt-relaxed-constexpr",
                "-shared",
                "-o",
                &lib,
            ])
            .args(
                ["iris_kernels.cu", "iris_abi.cu", "iris_batch.cu", "iris_reduce.cu", "iris_maskscan4.cu", "iris_dotbatch.cu", "iris_cluster.cu"]
                    .map(|f| format!("{src}/{f}")),
            )
            .args(["-ldl", "-lpthread"])
            .status()
            .expect("nvcc not found");
        assert!(status.success(), "nvcc failed");
        println!("cargo:rustc-link-search=native={out}");
        println!("cargo:rustc-link-lib=dylib=iris_b200");
        println!("cargo:rerun-if-changed={src}");
    }
    shadow_rs::new()
}
