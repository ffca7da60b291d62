// Link directives only.  MIRA_B200_LIB_DIR = the directory holding libmira_b200.so (mira_b200/ of the mira-b200
// repository after `make -C mira_b200/csrc`).  There is no CPU fallback to build: without the library the feature
// `b200` of the parent crate must stay off.
fn main() {
    println!("cargo:rerun-if-env-changed=MIRA_B200_LIB_DIR");
    let dir = std::env::var("MIRA_B200_LIB_DIR")
        .expect("set MIRA_B200_LIB_DIR to the directory that holds libmira_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=mira_b200");
    // let `cargo run` / `cargo test` find the library without LD_LIBRARY_PATH
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
}
