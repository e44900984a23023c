fn main() {
    // libvcb200.so is built by `make -C visioncortex_b200/csrc` (nvcc, sm_100a)
    println!("cargo:rustc-link-search=native=../visioncortex_b200/csrc");
    println!("cargo:rustc-link-lib=dylib=vcb200");
}
