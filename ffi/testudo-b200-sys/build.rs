// Links the in-tree shared library built by `python -m testudo_b200.build` (nvcc, sm_100a). The library has no CPU path:
// tb200_init fails without a CUDA device.
fn main() {
    let dir = std::env::var("TESTUDO_B200_LIB_DIR").unwrap_or_else(|_| "../../testudo_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=testudo_b200");
    println!("cargo:rerun-if-env-changed=TESTUDO_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/testudo_b200.h");
}
