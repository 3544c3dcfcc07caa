// Links against the prebuilt libqpzk.so (built by `make -C qp-zk-circuits_b200`).
fn main() {
    let dir = std::env::var("QPZK_LIB_DIR").unwrap_or_else(|_| "../../qp-zk-circuits_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=qpzk");
    println!("cargo:rerun-if-env-changed=QPZK_LIB_DIR");
}
