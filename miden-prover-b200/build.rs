// Link against libmiden_b200.so (built by `make -C miden-vm_b200/csrc` of the backend repository).
// MIDEN_B200_LIB_DIR points at the directory that holds it; the CUDA runtime it needs is resolved by the loader.
fn main() {
    println!("cargo:rerun-if-env-changed=MIDEN_B200_LIB_DIR");
    if let Ok(dir) = std::env::var("MIDEN_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=miden_b200");
}
