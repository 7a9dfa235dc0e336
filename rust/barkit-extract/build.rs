// Links libbarkit_b200.so.  BARKIT_B200_LIB_DIR = the directory that holds it (this repository:
// barkit_b200/, after `python -m barkit_b200.build`).
fn main() {
    let dir = std::env::var("BARKIT_B200_LIB_DIR")
        .expect("set BARKIT_B200_LIB_DIR to the directory containing libbarkit_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=barkit_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=BARKIT_B200_LIB_DIR");
}
