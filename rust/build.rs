// bindgen over the C-ABI header; links the in-tree shared library.
fn main() {
    let root = std::path::PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    println!("cargo:rustc-link-search=native={}", root.join("halo2-dynamic-sha256_b200").display());
    println!("cargo:rustc-link-lib=dylib=h2sha_b200");
    let bindings = bindgen::Builder::default()
        .header(root.join("include/h2sha_b200.h").to_str().unwrap())
        .allowlist_function("h2sha_.*")
        .allowlist_type("h2sha_.*")
        .allowlist_var("H2SHA_.*")
        .generate()
        .expect("bindgen");
    bindings
        .write_to_file(std::path::PathBuf::from(std::env::var("OUT_DIR").unwrap()).join("bindings.rs"))
        .unwrap();
}
