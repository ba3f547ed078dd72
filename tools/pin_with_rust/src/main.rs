//! Runs every committed golden vector of this repository through the REAL reference and writes what it produced to
//! `<golden dir>/rust_pins.json`; `tests/test_rust_pins.py` then holds the oracle, the engine, the regex-automata wire
//! reader and the ABI packer to those outputs.  One run on any machine with a Rust toolchain turns "parity unpinned"
//! into pinned.
//!
//!   emails_v1.json, rfc_vectors.json  -> zkemail_core::verify_email (core/src/circuits.rs:9-29; a panic is recorded with
//!                                        its message), cfdkim::canonicalize_signed_email (core/src/circuits.rs:34-35),
//!                                        VerificationOutput::abi_encode (core/src/io.rs:27-53)
//!   regex_v1.json (+ regex_pin_extra.json) -> dfa::regex::Regex::new + to_bytes_little_endian exactly as
//!                                        helpers/src/regex.rs:7-14 (create_dfa) and find_iter spans (core/src/regex.rs:36)
use base64::{engine::general_purpose::STANDARD as B64, Engine};
use regex_automata::dfa::regex::Regex as DFARegex;
use serde_json::{json, Value};
use std::panic::{catch_unwind, AssertUnwindSafe};
use zkemail_core::{verify_email, Email, PublicKey, VerificationOutput};

fn b64(v: &Value) -> Vec<u8> {
    B64.decode(v.as_str().expect("base64 string")).expect("valid base64")
}

fn panic_text(e: Box<dyn std::any::Any + Send>) -> String {
    if let Some(s) = e.downcast_ref::<&str>() {
        s.to_string()
    } else if let Some(s) = e.downcast_ref::<String>() {
        s.clone()
    } else {
        "panic".to_string()
    }
}

fn pin_email(g: &Value) -> Value {
    let raw = b64(&g["raw_email"]);
    let email = Email {
        from_domain: g["from_domain"].as_str().unwrap().to_string(),
        raw_email: raw.clone(),
        public_key: PublicKey { key: b64(&g["key"]), key_type: g["key_type"].as_str().unwrap().to_string() },
        external_inputs: vec![],
    };
    let mut out = json!({ "name": g["name"] });
    match catch_unwind(AssertUnwindSafe(|| verify_email(&email))) {
        Ok(v) => {
            out["status"] = json!("ok");
            out["from_domain_hash"] = json!(hex::encode(&v.from_domain_hash));
            out["public_key_hash"] = json!(hex::encode(&v.public_key_hash));
            out["abi_encode"] = json!(hex::encode(VerificationOutput::from_parts(v, None).abi_encode()));
        }
        Err(e) => {
            out["status"] = json!("panic");
            out["panic_message"] = json!(panic_text(e));
        }
    }
    // the regex haystacks: header preimage and canonical body of the first valid DKIM-Signature header
    let canon = catch_unwind(AssertUnwindSafe(|| {
        let parsed = mailparse::parse_mail(&raw).unwrap();
        let logger = slog::Logger::root(slog::Discard, slog::o!());
        cfdkim::canonicalize_signed_email(&logger, &parsed).unwrap()
    }));
    match canon {
        Ok((header, body, signature)) => {
            out["canon_header"] = json!(B64.encode(&header));
            out["canon_body"] = json!(B64.encode(&body));
            out["signature"] = json!(B64.encode(&signature));
        }
        Err(e) => out["canon_error"] = json!(panic_text(e)),
    }
    out
}

fn pin_regex(pattern: &str, haystacks: &[Vec<u8>]) -> Value {
    match DFARegex::new(pattern) {
        Err(e) => json!({ "pattern": pattern, "error": e.to_string() }),
        Ok(re) => {
            // helpers/src/regex.rs:7-14
            let (fwd, fwd_pad) = re.forward().to_bytes_little_endian();
            let (bwd, bwd_pad) = re.reverse().to_bytes_little_endian();
            let hs: Vec<Value> = haystacks
                .iter()
                .map(|h| {
                    let spans: Vec<[usize; 2]> = re.find_iter(h).map(|m| [m.start(), m.end()]).collect();
                    json!({ "haystack": B64.encode(h), "spans": spans })
                })
                .collect();
            json!({ "pattern": pattern, "fwd": B64.encode(&fwd[fwd_pad..]), "bwd": B64.encode(&bwd[bwd_pad..]), "haystacks": hs })
        }
    }
}

fn main() {
    let dir = std::env::args().nth(1).expect("usage: pin_with_rust <tests/golden directory>");
    let read = |name: &str| -> Value {
        serde_json::from_str(&std::fs::read_to_string(format!("{dir}/{name}")).expect(name)).expect("json")
    };
    let mut emails = vec![];
    for file in ["emails_v1.json", "rfc_vectors.json"] {
        for g in read(file).as_array().unwrap() {
            emails.push(pin_email(g));
        }
    }
    // regex_v1.json: [{pattern, haystack, spans}] -> group the haystacks by pattern
    let mut by_pattern: Vec<(String, Vec<Vec<u8>>)> = vec![];
    // regex_pin_extra.json: [{pattern, haystack}] without expected spans - patterns on which backtracking engines and the
    // Thompson construction may disagree (loops whose body can match the empty string); the crate's answer is the pin
    let mut groups: Vec<Value> = read("regex_v1.json").as_array().unwrap().clone();
    if std::path::Path::new(&format!("{dir}/regex_pin_extra.json")).exists() {
        groups.extend(read("regex_pin_extra.json").as_array().unwrap().iter().cloned());
    }
    for g in groups.iter() {
        let p = g["pattern"].as_str().unwrap().to_string();
        let h = b64(&g["haystack"]);
        match by_pattern.iter_mut().find(|(q, _)| *q == p) {
            Some((_, hs)) => hs.push(h),
            None => by_pattern.push((p, vec![h])),
        }
    }
    let regex: Vec<Value> = by_pattern.iter().map(|(p, hs)| pin_regex(p, hs)).collect();
    let out = json!({
        "reference": "zkemail/zkemail.rs (zkemail-core) with the dependency set of its Cargo.lock",
        "emails": emails,
        "regex": regex,
    });
    std::fs::write(format!("{dir}/rust_pins.json"), serde_json::to_string_pretty(&out).unwrap()).unwrap();
    println!("wrote {dir}/rust_pins.json: {} emails, {} patterns", out["emails"].as_array().unwrap().len(), regex.len());
}
