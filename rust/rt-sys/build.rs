// UNCOMPILED in this repository's image: see rust/README.md.
fn main() {
    // librt_b200.so is built by `make -C raytracing-course-2024_b200/csrc` (nvcc, sm_100a)
    let dir = std::env::var("RT_B200_LIB_DIR").expect("set RT_B200_LIB_DIR to .../raytracing-course-2024_b200/_build");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=rt_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=RT_B200_LIB_DIR");
}
