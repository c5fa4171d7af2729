// Tells rustc where libpgvrx_b200.so lives.  The library is built by `python -m pgvector_rx_b200.build` (nvcc, sm_100a);
// point PGVRX_B200_LIB_DIR at the directory that holds it (default: ../../pgvector_rx_b200 relative to this crate).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("PGVRX_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../pgvector_rx_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=pgvrx_b200");
    // the library carries its CUDA runtime statically; the loader still needs to find the .so itself at run time
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=PGVRX_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/pgvrx_b200.h");
}
