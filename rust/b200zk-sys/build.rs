// cc-built .cu, extern "C" launchers, one architecture, no fallback (BASELINE.json north_star).
// Set B200ZK_LIB_DIR to link a prebuilt libb200zk.so (zk-apps_b200/build.py) instead of compiling.
use std::{env, fs, path::PathBuf};

fn main() {
    if let Ok(dir) = env::var("B200ZK_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=b200zk");
        return;
    }
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("zk-apps_b200/csrc");
    let mut b = cc::Build::new();
    b.cuda(true)
        .cudart("static")
        .flag("-std=c++17")
        .flag("-O3")
        .flag("-lineinfo")
        .flag("--expt-relaxed-constexpr")
        .flag("-gencode")
        .flag("arch=compute_100a,code=sm_100a")
        .include(root.join("include"))
        .include(&csrc);
    for e in fs::read_dir(&csrc).unwrap() {
        let p = e.unwrap().path();
        println!("cargo:rerun-if-changed={}", p.display());
        if p.extension().map_or(false, |x| x == "cu") {
            b.file(&p);
        }
    }
    println!("cargo:rerun-if-changed={}", root.join("include/b200zk.h").display());
    b.compile("b200zk");
}
