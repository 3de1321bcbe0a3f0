// Links libsalg_b200.so (needs libcudart and libnccl at run time).  SALG_B200_LIB_DIR = directory holding the library
// (e.g. <repo>/single-algebra_b200 after `python -c "import __graft_entry__ as g; g.build()"`).
fn main() {
    let dir = std::env::var("SALG_B200_LIB_DIR").expect("set SALG_B200_LIB_DIR to the directory of libsalg_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=salg_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=SALG_B200_LIB_DIR");
}
