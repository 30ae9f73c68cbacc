// SOURCE ONLY (never compiled in this image).  Point KDNB_LIB_DIR at the directory holding libkdnb.so; the default is
// multilanguagekdtree_b200/ of this repository, resolved from the crate's own location.
use std::path::PathBuf;

fn main() {
    let dir = match std::env::var("KDNB_LIB_DIR") {
        Ok(d) => PathBuf::from(d),
        Err(_) => PathBuf::from(env!("CARGO_MANIFEST_DIR")).join("../../multilanguagekdtree_b200"),
    };
    let dir = dir.canonicalize().unwrap_or(dir);
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=kdnb");
    println!("cargo:rerun-if-env-changed=KDNB_LIB_DIR");
}
