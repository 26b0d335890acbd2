// Link against the in-tree library built by `python -c "import __graft_entry__ as g; g.build()"`.
fn main() {
    let dir = std::env::var("BEVY_GGRS_B200_LIB_DIR").unwrap_or_else(|_| "../../bevy_ggrs_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=bevy_ggrs_b200");
}
