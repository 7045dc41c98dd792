// Links libgranne_b200.so (built by `python -m granne_b200.build`); set GRANNE_B200_LIB_DIR to its directory.
fn main() {
    if let Ok(dir) = std::env::var("GRANNE_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=granne_b200");
    println!("cargo:rerun-if-env-changed=GRANNE_B200_LIB_DIR");
}
