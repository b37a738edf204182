// Links libsliceslice_b200.so the way the reference links its C++ competitor
// (bench/sse4-strstr/build.rs:16-23), minus cc/bindgen: the library is prebuilt by
// `python -m sliceslice_rs_b200.build` and the bindings below are written by hand.
fn main() {
    let dir = std::env::var("SLICESLICE_B200_LIB_DIR").expect("set SLICESLICE_B200_LIB_DIR to the directory holding libsliceslice_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=sliceslice_b200");
    println!("cargo:rerun-if-env-changed=SLICESLICE_B200_LIB_DIR");
}
