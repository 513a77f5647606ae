// build.rs -- only does something with `--features b200`: tells rustc where libringzk_b200.so lives.
// RINGZK_B200_LIB_DIR = the directory holding the library built by the engine's repository
// (python -c "import __graft_entry__ as g; g.build()"  ->  ring-zk_b200/_build/libringzk_b200.so).
fn main() {
    println!("cargo:rerun-if-env-changed=RINGZK_B200_LIB_DIR");
    if std::env::var("CARGO_FEATURE_B200").is_ok() {
        let dir = std::env::var("RINGZK_B200_LIB_DIR").expect("set RINGZK_B200_LIB_DIR to the directory of libringzk_b200.so");
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-lib=dylib=ringzk_b200");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
}
