// rtc-sys/build.rs — builds librtc_b200.so from the CUDA / C++ sources for sm_100a.
//
// The list of translation units and their flags is NOT repeated here: it is read from csrc/manifest.txt, the same file
// ray-tracer-challenge-rust_b200/build.py reads, so the two builds cannot drift apart.
//
//   RTC_B200_CSRC   directory holding manifest.txt and the sources (default: ../ray-tracer-challenge-rust_b200/csrc
//                   relative to this crate)
//   NVCC / CXX      compilers (default: nvcc, g++)
use std::env;
use std::fs;
use std::path::{Path, PathBuf};
use std::process::Command;

fn run(cmd: &mut Command) {
    let status = cmd.status().unwrap_or_else(|e| panic!("cannot start {:?}: {}", cmd, e));
    assert!(status.success(), "{:?} failed", cmd);
}

/// The X(...) entries of `#define <name>(X) X(1) X(2) ...` in a header.
fn instance_masks(header: &Path, name: &str) -> Vec<String> {
    let text = fs::read_to_string(header).unwrap_or_else(|e| panic!("{}: {}", header.display(), e));
    let key = format!("#define {}(X)", name);
    let line = text.lines().find(|l| l.trim_start().starts_with(&key)).unwrap_or_else(|| panic!("{} not found", key));
    line[line.find(&key).unwrap() + key.len()..]
        .split("X(")
        .skip(1)
        .map(|s| s[..s.find(')').expect("unbalanced X(")].trim().to_string())
        .collect()
}

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let here = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = env::var("RTC_B200_CSRC").map(PathBuf::from).unwrap_or_else(|_| here.join("../ray-tracer-challenge-rust_b200/csrc"));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let cxx = env::var("CXX").unwrap_or_else(|_| "g++".into());
    let manifest = csrc.join("manifest.txt");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-env-changed=RTC_B200_CSRC");

    let text = fs::read_to_string(&manifest).unwrap_or_else(|e| panic!("{}: {}", manifest.display(), e));
    let mut hostflags: Vec<String> = vec![];
    let mut nvccflags: Vec<String> = vec![];
    let mut linkflags: Vec<String> = vec![];
    let mut objects: Vec<PathBuf> = vec![];
    // (compiler, flags, source, object, extra define)
    let mut jobs: Vec<(bool, PathBuf, PathBuf, Option<String>)> = vec![];
    for line in text.lines() {
        let line = line.trim();
        if line.is_empty() || line.starts_with('#') {
            continue;
        }
        let mut words = line.split_whitespace();
        match words.next().unwrap() {
            "hostflags" => hostflags = words.map(String::from).collect(),
            "nvccflags" => nvccflags = words.map(String::from).collect(),
            "link" => linkflags.extend(words.map(String::from)),
            kind @ ("host" | "cuda") => {
                let file = words.next().expect("file name");
                let obj = out.join(format!("{}.o", file.replace('.', "_")));
                jobs.push((kind == "cuda", csrc.join(file), obj, None));
            }
            "instances" => {
                let file = words.next().expect("file name");
                let define = words.next().expect("macro name");
                let (header, list) = words.next().expect("header:macro").split_once(':').expect("header:macro");
                for mask in instance_masks(&csrc.join(header), list) {
                    let obj = out.join(format!("{}_{}.o", file.replace('.', "_"), mask));
                    jobs.push((true, csrc.join(file), obj, Some(format!("-D{}={}", define, mask))));
                }
            }
            other => panic!("manifest.txt: unknown directive {}", other),
        }
    }
    // translation units are independent: compile them on scoped threads
    let handles: Vec<_> = jobs
        .iter()
        .cloned()
        .map(|(cuda, src, obj, define)| {
            let (compiler, flags) = if cuda { (nvcc.clone(), nvccflags.clone()) } else { (cxx.clone(), hostflags.clone()) };
            std::thread::spawn(move || {
                let mut c = Command::new(&compiler);
                c.args(&flags);
                if let Some(d) = &define {
                    c.arg(d);
                }
                c.arg("-c").arg(&src).arg("-o").arg(&obj);
                run(&mut c);
            })
        })
        .collect();
    for h in handles {
        h.join().expect("compile job");
    }
    for (_, _, obj, _) in &jobs {
        objects.push(obj.clone());
    }
    let lib = out.join("librtc_b200.so");
    run(Command::new(&nvcc).arg("-shared").args(&linkflags).arg("-o").arg(&lib).args(&objects));
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=rtc_b200");
    // so that `cargo run` finds the library without LD_LIBRARY_PATH
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
}
