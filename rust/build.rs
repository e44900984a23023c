// Links the CUDA product library.  VCB200_LIB_DIR = directory holding libvcb200.so (visioncortex_b200/csrc after `make`).
fn main() {
    let dir = std::env::var("VCB200_LIB_DIR").unwrap_or_else(|_| "../visioncortex_b200/csrc".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=vcb200");
    println!("cargo:rerun-if-env-changed=VCB200_LIB_DIR");
}
