//! Writes REFERENCE test vectors: runs the reference's own CPU path (omr_core + Primus-fhe) with a fixed seed and dumps
//! keys, clues, every stage boundary of `detect`, the pertinency vector and a payload digest as `*.omrb` blobs.
//!
//!     cargo run --release --example dump_vectors -- <out_dir> [n_messages = 4] [seed = 20261018]
//!     cp <out_dir>/*.omrb  <repo>/tests/golden/ref_v1/      # then: python -m pytest tests/test_ref_vectors.py
//!
//! `tests/test_ref_vectors.py` makes the CUDA library and the CPU oracle reproduce every file word for word; it is skipped
//! while the directory is empty.  This is what pins bit-level parity with Primus-fhe ("parity unpinned").
//!
//! The stage sequence is the public-API one of `omr_core/benches/two_level_bs.rs:12-145` (the stage functions of detector.rs
//! are private); the final ciphertext is cross-checked against `Detector::detect` itself before anything is written.
//! All ring polynomials are written in coefficient form (domain = 1).

use std::{env, path::PathBuf};

use algebra::{
    ntt::NumberTheoryTransform,
    reduce::{ModulusValue, Reduce, ReduceAddAssign},
    Field,
};
use fhe_core::{lwe_modulus_switch, lwe_modulus_switch_assign, LweCiphertext, RlweCiphertext};
use omr_core::{ClueValue, FirstLevelField, InterLweValue, KeyGen, OmrParameters, Payload, SecondLevelField};
use omr_b200_sys::{blob, flatten, sys};
use rand::{rngs::StdRng, Rng, SeedableRng};

fn main() {
    let args: Vec<String> = env::args().collect();
    let out = PathBuf::from(args.get(1).expect("usage: dump_vectors <out_dir> [n_messages] [seed]"));
    let n: usize = args.get(2).map(|s| s.parse().unwrap()).unwrap_or(4);
    let seed: u64 = args.get(3).map(|s| s.parse().unwrap()).unwrap_or(20261018);
    std::fs::create_dir_all(&out).unwrap();
    let mut rng = StdRng::seed_from_u64(seed);

    let params = OmrParameters::new();
    let secret_key_pack = KeyGen::generate_secret_key(params.clone(), &mut rng);
    let decoy_pack = KeyGen::generate_secret_key(params.clone(), &mut rng);
    let sender = secret_key_pack.generate_sender(&mut rng);
    let decoy = decoy_pack.generate_sender(&mut rng);
    let detector = secret_key_pack.generate_detector(&mut rng);
    let detection_key = detector.detection_key();
    let table2 = secret_key_pack.second_level_ntt_table();

    // ---- keys ------------------------------------------------------------------------------------------------------------
    let flat = flatten::flatten_detection_key(detection_key);
    blob::write(&out.join("detection_key.omrb"), blob::DETECTION_KEY, 0, 0, 0, blob::DOMAIN_COEFF,
                &[blob::bytes_of(&flat.bsk1), blob::bytes_of(&flat.ksk), blob::bytes_of(&flat.bsk2), blob::bytes_of(&flat.trace)]).unwrap();
    // secrets as i32 (binary / ternary; ternary -1 is stored as q - 1 by the reference: secret.rs:134-138)  [UPSTREAM: as_ref()]
    let lift = |v: u64, q: u64| if v == q - 1 { -1i32 } else { v as i32 };
    let s0: Vec<i32> = secret_key_pack.clue_secret_key().as_ref().iter().map(|&v| v as i32).collect();
    let z1: Vec<i32> = secret_key_pack.first_level_rlwe_secret_key().as_ref().iter().map(|&v| lift(v as u64, sys::OMR_Q1 as u64)).collect();
    let s2: Vec<i32> = secret_key_pack.intermediate_lwe_secret_key().as_ref().iter().map(|&v| v as i32).collect();
    let z2: Vec<i32> = secret_key_pack.second_level_rlwe_secret_key().as_ref().iter().map(|&v| lift(v, sys::OMR_Q2)).collect();
    blob::write(&out.join("secret_key.omrb"), blob::SECRET_KEY, 0, 0, 0, 0,
                &[blob::bytes_of(&s0), blob::bytes_of(&z1), blob::bytes_of(&s2), blob::bytes_of(&z2)]).unwrap();

    // ---- clues: messages 0, 2, 4, ... pertinent, the others under the decoy key (examples/omr.rs:126-135) -----------------------
    let clues: Vec<_> = (0..n).map(|i| if i % 2 == 0 { sender.gen_clues(&mut rng) } else { decoy.gen_clues(&mut rng) }).collect();
    let (ca, cb) = flatten::flatten_clues(&clues);
    blob::write(&out.join("clues.omrb"), blob::CLUES, n as u64, 0, 0, 0, &[blob::bytes_of(&ca), blob::bytes_of(&cb)]).unwrap();

    // ---- stage boundaries, per message (benches/two_level_bs.rs:24-145) -----------------------------------------------------
    let mut rlwe1 = Vec::<u32>::new();
    let mut lwe2 = Vec::<u32>::new();
    let mut rlwe2 = Vec::<u64>::new();
    let mut pv = Vec::<u64>::new();
    let mut pertinency_vector = Vec::new();
    let clue_modulus_value = params.clue_params().cipher_modulus_value;
    let n1 = params.first_level_ring_dimension();
    let inter = params.intermediate_lwe_params();
    for clue in &clues {
        let mut lwes: Vec<LweCiphertext<ClueValue>> = clue.extract_all(detection_key.clue_modulus());
        if clue_modulus_value != ModulusValue::PowerOf2(n1 as ClueValue * 2) {
            lwes.iter_mut().for_each(|c| lwe_modulus_switch_assign(c, clue_modulus_value, n1 as ClueValue * 2));
        }
        let brk1 = detection_key.first_level_blind_rotation_key();
        let sum = lwes.iter().map(|c| brk1.blind_rotate(detector.first_level_lut().clone(), c))
            .reduce(|acc, e| acc.add_element_wise(&e))
            .unwrap_or_else(|| <RlweCiphertext<FirstLevelField>>::zero(n1));
        rlwe1.extend(sum.a().iter().copied());
        rlwe1.extend(sum.b().iter().copied());
        let ks = detection_key.first_level_key_switching_key().key_switch(&sum.extract_lwe_locally(), FirstLevelField::MODULUS);
        let mut inter_lwe = lwe_modulus_switch(&ks, params.first_level_blind_rotation_params().modulus, inter.cipher_modulus_value);
        let log_plain = inter.plain_modulus_value.trailing_zeros();
        let scale = (clue.msg_count() as InterLweValue) * match inter.cipher_modulus_value {
            ModulusValue::Native => 1 << (InterLweValue::BITS - log_plain),
            ModulusValue::PowerOf2(q) => q >> log_plain,
            ModulusValue::Prime(q) | ModulusValue::Others(q) => ((q >> (log_plain - 1)) + 1) >> 1,
        };
        inter.cipher_modulus.reduce_add_assign(inter_lwe.b_mut(), inter.cipher_modulus.reduce(scale));
        if inter.cipher_modulus_value != ModulusValue::PowerOf2(params.second_level_ring_dimension() as InterLweValue * 2) {
            lwe_modulus_switch_assign(&mut inter_lwe, inter.cipher_modulus_value, params.second_level_ring_dimension() as InterLweValue * 2);
        }
        lwe2.extend(inter_lwe.a().iter().copied());
        lwe2.push(inter_lwe.b());
        let mut second = detection_key.second_level_blind_rotation_key().blind_rotate(detector.second_level_lut().clone(), &inter_lwe);
        rlwe2.extend(second.a().iter().copied());
        rlwe2.extend(second.b().iter().copied());
        let n_inv = detection_key.second_level_ring_dimension_inv();
        second.a_mut().mul_shoup_scalar_assign(n_inv);
        second.b_mut().mul_shoup_scalar_assign(n_inv);
        let staged = detection_key.trace_key().trace(&second).to_ntt_rlwe(table2);
        // the public API must agree with the stage sequence above, word for word
        let whole = detector.detect(clue);
        assert!(staged.a() == whole.a() && staged.b() == whole.b(), "stage sequence differs from Detector::detect");
        flatten::ntt_rlwe_to_coeff(&whole, table2, &mut pv);
        pertinency_vector.push(whole);
    }
    blob::write(&out.join("rlwe1.omrb"), blob::RLWE1, n as u64, 0, 0, blob::DOMAIN_COEFF, &[blob::bytes_of(&rlwe1)]).unwrap();
    blob::write(&out.join("lwe2.omrb"), blob::LWE2, n as u64, 0, 0, 0, &[blob::bytes_of(&lwe2)]).unwrap();
    blob::write(&out.join("rlwe2.omrb"), blob::RLWE2, n as u64, 0, 0, blob::DOMAIN_COEFF, &[blob::bytes_of(&rlwe2)]).unwrap();
    blob::write(&out.join("pertinency_vector.omrb"), blob::PERTINENCY_VECTOR, n as u64, 0, 0, blob::DOMAIN_COEFF, &[blob::bytes_of(&pv)]).unwrap();

    // ---- payload digest (deterministic given the rng seed; the index digest uses thread_rng buckets and is not) --------------
    let payloads: Vec<Payload> = (0..n).map(|_| Payload::random(&mut rng)).collect();
    let flat_payloads: Vec<u16> = payloads.iter().flat_map(|p| p.0.iter().copied()).collect();
    blob::write(&out.join("payloads.omrb"), blob::PAYLOADS, n as u64, 0, 0, 0, &[blob::bytes_of(&flat_payloads)]).unwrap();
    let pertinent = n.div_ceil(2);
    let retriever = secret_key_pack.generate_retriever(n, pertinent);
    let rp = retriever.params();
    let weight_seed = [seed as u8; 32]; // 32-byte StdRng seed: every byte = low byte of `seed` (aux of the digest blob)
    let digest = detector.encode_pertinent_payloads(&pertinency_vector, &payloads, rp.combination_count(), rp.cmb_count_per_cipher(),
                                                    &mut StdRng::from_seed(weight_seed));
    let mut flat_digest = Vec::<u64>::new();
    for ct in &digest {
        flatten::ntt_rlwe_to_coeff(ct, table2, &mut flat_digest);
    }
    blob::write(&out.join("payload_digest.omrb"), blob::DIGEST, digest.len() as u64, 0, (seed as u8) as u64, blob::DOMAIN_COEFF, &[blob::bytes_of(&flat_digest)]).unwrap();
    let _: u8 = rng.gen();
    println!("wrote {} messages to {}", n, out.display());
}
