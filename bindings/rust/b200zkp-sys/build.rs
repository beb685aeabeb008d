// Links the in-tree shared library.  B200ZKP_LIB_DIR overrides the default (<repo>/boundless_b200).
fn main() {
    let dir = std::env::var("B200ZKP_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{manifest}/../../../boundless_b200")
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=b200zkp");
    println!("cargo:rerun-if-env-changed=B200ZKP_LIB_DIR");
}
