// Locates libbfgpu.so: $BFGPU_LIB_DIR (the directory holding the library built by `python -c 'import __graft_entry__ as g; g.build()'`,
// i.e. <backend checkout>/zkvm-brainfuck_b200) or the system library path.
fn main() {
    println!("cargo:rerun-if-env-changed=BFGPU_LIB_DIR");
    if let Ok(dir) = std::env::var("BFGPU_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=bfgpu");
}
