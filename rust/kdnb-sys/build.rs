// SOURCE ONLY (never compiled in this image).  Point KDNB_LIB_DIR at the directory holding libkdnb.so
// (multilanguagekdtree_b200/ in this repository).
fn main() {
    let dir = std::env::var("KDNB_LIB_DIR").unwrap_or_else(|_| "../../multilanguagekdtree_b200".into());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=kdnb");
    println!("cargo:rerun-if-env-changed=KDNB_LIB_DIR");
}
