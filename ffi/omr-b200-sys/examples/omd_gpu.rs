//! `omr_core/examples/omd.rs` with the detector swapped for the GPU one: one pertinent and one non-pertinent clue, detect on
//! the B200, decrypt with the reference's own secret key and NTT table, same assertions (omd.rs:52-58).
//!
//!     cargo run --release --example omd_gpu

use algebra::{ntt::NumberTheoryTransform, Field};
use omr_core::{KeyGen, OmrParameters, SecondLevelField};
use omr_b200_sys::GpuDetector;

type Inner = <SecondLevelField as Field>::ValueT;

fn main() {
    let params = OmrParameters::new();
    let mut rng = rand::thread_rng();
    let fp = <SecondLevelField as Field>::MODULUS_VALUE;
    let ft = params.output_plain_modulus_value();
    let decode = |c: Inner| (c as f64 * ft as f64 / fp as f64).round() as Inner % ft;

    let secret_key_pack = KeyGen::generate_secret_key(params.clone(), &mut rng);
    let secret_key_pack2 = KeyGen::generate_secret_key(params.clone(), &mut rng);
    let key = secret_key_pack.second_level_ntt_rlwe_secret_key();
    let ntt_table = secret_key_pack.second_level_ntt_table();
    let sender = secret_key_pack.generate_sender(&mut rng);
    let sender2 = secret_key_pack2.generate_sender(&mut rng);
    let detector = GpuDetector::new(secret_key_pack.generate_detection_key(&mut rng)); // the only changed line

    let clues = sender.gen_clues(&mut rng);
    let clues2 = sender2.gen_clues(&mut rng);
    let mut results = detector.detect_batch(&[clues, clues2]);
    let result2 = results.pop().unwrap();
    let result = results.pop().unwrap();

    let poly = ntt_table.inverse_transform_inplace(result.b() - result.a().clone() * &**key);
    let decrypted = poly.into_iter().map(decode).collect::<Vec<Inner>>();
    assert_eq!(decrypted[0], 1);
    assert!(decrypted[1..].iter().all(|&x| x == 0));
    let poly2 = ntt_table.inverse_transform_inplace(result2.b() - result2.a().clone() * &**key);
    assert!(poly2.into_iter().map(decode).all(|x| x == 0));
    println!("omd on the GPU: ok");
}
