// Links the prebuilt librtc_b200.so (built by `python -m ray_tracer_challenge_b200.build`, nvcc sm_100a).
// RTC_B200_LIB_DIR points at the directory that holds it.
fn main() {
    let dir = std::env::var("RTC_B200_LIB_DIR").unwrap_or_else(|_| "../../ray_tracer_challenge_b200".into());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rtc_b200");
    println!("cargo:rerun-if-env-changed=RTC_B200_LIB_DIR");
}
