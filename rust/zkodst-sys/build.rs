// Links against the in-tree shared library built by `python zk-odst_b200/build.py`.
fn main() {
    let dir = std::env::var("ZKODST_LIB_DIR").unwrap_or_else(|_| "../../zk-odst_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=zkodst");
}
