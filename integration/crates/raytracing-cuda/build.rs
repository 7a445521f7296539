// bindgen over include/rtcuda.h, the way crates/raytracing-optix/build.rs:3-55 runs it over csrc/host/lib_api.h.
// RTCUDA_DIR = a checkout of the backend: $RTCUDA_DIR/include/rtcuda.h and
// $RTCUDA_DIR/opencl-raytracing_b200/libraytracing_cuda.so (built by `make`, nvcc -gencode arch=compute_100a,code=sm_100a).
use std::{env, path::PathBuf};

fn main() {
    let dir = PathBuf::from(env::var("RTCUDA_DIR").expect("set RTCUDA_DIR to the raytracing-cuda backend checkout"));
    let lib_dir = dir.join("opencl-raytracing_b200");
    println!("cargo:rerun-if-env-changed=RTCUDA_DIR");
    println!("cargo:rerun-if-changed={}", dir.join("include/rtcuda.h").display());
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=raytracing_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib_dir.display());

    let bindings = bindgen::Builder::default()
        .header(dir.join("include/rtcuda.h").to_string_lossy())
        .allowlist_function("rtcuda_.*")
        .allowlist_type("rtcuda_.*")
        .allowlist_var("RTCUDA_.*")
        .rustified_enum("rtcuda_status")
        .derive_default(true)
        .generate()
        .expect("bindgen over rtcuda.h");
    bindings
        .write_to_file(PathBuf::from(env::var("OUT_DIR").unwrap()).join("bindings.rs"))
        .expect("write bindings.rs");
}
