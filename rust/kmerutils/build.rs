// Links the shim against libkmerutils_b200.so (built by `python -m kmerutils_b200.build`, nvcc sm_100a).
//   KMERUTILS_B200_DIR = directory holding libkmerutils_b200.so (default: ../../kmerutils_b200 relative to this crate)
fn main() {
    let dir = std::env::var("KMERUTILS_B200_DIR").unwrap_or_else(|_| {
        let here = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{here}/../../kmerutils_b200")
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=kmerutils_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=KMERUTILS_B200_DIR");
    println!("cargo:rerun-if-changed=build.rs");
}
