// Links libb2f.so.  B2F_LIB_DIR = directory that holds it (default: ../../libflate_b200 of this repository).
fn main() {
    let dir = std::env::var("B2F_LIB_DIR").unwrap_or_else(|_| {
        let here = std::path::PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap());
        here.join("../../libflate_b200").to_string_lossy().into_owned()
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=b2f");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=B2F_LIB_DIR");
}
