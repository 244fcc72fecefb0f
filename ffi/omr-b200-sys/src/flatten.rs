//! Flattening of the reference's (Primus-fhe) types into the flat arrays of `include/omr_b200.h`, and back.
//!
//! Everything crosses the boundary in **coefficient form**: NTT-domain polynomials are passed through Primus-fhe's own table
//! (`inverse_transform` on the way out of Rust, `transform` on the way in), so the C library never needs to agree with
//! Primus-fhe on the root of unity or on the order of NTT-domain vectors (risk 1).
//!
//! `[UPSTREAM]` marks every accessor of an un-vendored Primus-fhe type (`algebra`, `lattice`, `fhe_core`, branch `omr2`) whose
//! exact name could not be checked when this file was written — the six adapter functions right below are the only place to
//! adjust.  Layouts on the C side (row-major, little-endian):
//!
//! | Rust (key_gen/detection.rs:9-16)                          | flat                                                                  |
//! |---|---|
//! | `BlindRotationKey<FirstLevelField>`  = 512 x `NttRgsw`     | `bsk1 [512][8][2][1024] u32`: rows 0..3 = `c_neg_s_m` (RLWE(-z m g_j)), rows 4..7 = `c_m` (RLWE(m g_j)); `[.][.][0]` = a, `[1]` = b |
//! | `NonPowOf2LweKeySwitchingKey<u32>` = 1024 x 27 LWE         | `ksk [1024][27][671] u32`: `(a[670], b)` of LWE_{s2}(z1[i] 2^j)       |
//! | `BlindRotationKey<SecondLevelField>` = 670 x `NttRgsw`     | `bsk2 [670][12][2][2048] u64`: rows 0..5 = `c_neg_s_m`, 6..11 = `c_m` |
//! | `TraceKey<SecondLevelField>` = 11 x `NttGadgetRlwe` (25)   | `trace [11][25][2][2048] u64`: step t <-> automorphism X -> X^(2^(11-t)+1) (the order `TraceKey::trace` applies them) |
//!
//! Gadget row j of a key must encrypt `m * 2^(drop + w*j)` (first level: w = 5, 4 rows, drop 7; second level: w = 7, 6 rows,
//! drop 8; trace: w = 2, 25 rows, drop 0) — the convention of `NonPowOf2ApproxSignedBasis` as restated in SURVEY App. A.4.  If
//! Primus-fhe stores rows most-significant first, reverse them in `gadget_rows` below.

use algebra::{ntt::NumberTheoryTransform, Field, NttField};
use fhe_core::{BlindRotationKey, CmLweCiphertext, NonPowOf2LweKeySwitchingKey, NttRlweCiphertext, TraceKey};
use lattice::{NttGadgetRlwe, NttRgsw, NttRlwe};
use omr_core::{ClueValue, DetectionKey, FirstLevelField, SecondLevelField};

use crate::sys;

// ---- [UPSTREAM] adapters: the only code that names Primus-fhe accessors the reference itself does not use ---------------------

/// the RGSW ciphertexts of a blind-rotation key, in LWE-secret order  `[UPSTREAM] BlindRotationKey::key() -> &[NttRgsw<F>]`
fn rgsw_list<F: NttField>(key: &BlindRotationKey<F>) -> &[NttRgsw<F>] {
    key.key()
}
/// the two gadget halves of an RGSW ciphertext: (RLWE(-s m g_j))_j and (RLWE(m g_j))_j  `[UPSTREAM] NttRgsw::{minus_s_m, m}`
fn rgsw_halves<F: NttField>(rgsw: &NttRgsw<F>) -> (&NttGadgetRlwe<F>, &NttGadgetRlwe<F>) {
    (rgsw.minus_s_m(), rgsw.m())
}
/// the rows of a gadget RLWE, least-significant gadget power first  `[UPSTREAM] NttGadgetRlwe::data() -> &[NttRlwe<F>]`
fn gadget_rows<F: NttField>(g: &NttGadgetRlwe<F>) -> &[NttRlwe<F>] {
    g.data()
}
/// the LWE rows of the key-switching key: `[i][j]` -> (a, b) with a of the output dimension  `[UPSTREAM]`
fn ksk_row(ksk: &NonPowOf2LweKeySwitchingKey<u32>, i: usize, j: usize) -> (&[u32], u32) {
    let lwe = &ksk.key()[i][j];
    (lwe.a(), lwe.b())
}
/// the automorphism keys of the trace key in the order `TraceKey::trace` applies them  `[UPSTREAM] TraceKey::keys()`
fn trace_steps<F: NttField>(tk: &TraceKey<F>) -> Vec<&NttGadgetRlwe<F>> {
    tk.keys().iter().map(|auto_key| auto_key.key()).collect()
}
/// build an NTT-domain RLWE ciphertext from its two NTT-domain polynomials  `[UPSTREAM] NttRlwe::new(a, b)`
fn ntt_rlwe_from_parts<F: NttField>(a: Vec<<F as Field>::ValueT>, b: Vec<<F as Field>::ValueT>) -> NttRlwe<F> {
    NttRlwe::new(algebra::polynomial::FieldNttPolynomial::new(a), algebra::polynomial::FieldNttPolynomial::new(b))
}

// ---- keys -------------------------------------------------------------------------------------------------------------------

/// The four flat arrays of `omr_key_blobs`, coefficient form (`OMR_KEYS_COEFF`).
pub struct FlatDetectionKey {
    pub bsk1: Vec<u32>,
    pub ksk: Vec<u32>,
    pub bsk2: Vec<u64>,
    pub trace: Vec<u64>,
}

impl FlatDetectionKey {
    pub fn blobs(&self) -> sys::OmrKeyBlobs {
        sys::OmrKeyBlobs { bsk1: self.bsk1.as_ptr(), ksk: self.ksk.as_ptr(), bsk2: self.bsk2.as_ptr(), trace: self.trace.as_ptr(), flags: sys::OMR_KEYS_COEFF }
    }
}

/// one NTT-domain RLWE row -> `[a | b]` in coefficient form, appended to `out`
fn push_row_coeff<F: NttField>(row: &NttRlwe<F>, table: &<F as NttField>::Table, out: &mut Vec<<F as Field>::ValueT>) {
    for poly in [row.a(), row.b()] {
        let mut c = poly.clone();
        // the same inverse transform the reference applies to a detect result (examples/omd.rs:48)
        let coeff = table.inverse_transform_inplace(std::mem::take(&mut c));
        out.extend(coeff.into_iter());
    }
}

fn flatten_brk<F: NttField>(key: &BlindRotationKey<F>, levels: usize, n: usize) -> Vec<<F as Field>::ValueT> {
    let table = key.ntt_table(); // used by the reference itself: detector.rs:231-234
    let list = rgsw_list(key);
    let mut out = Vec::with_capacity(list.len() * 2 * levels * 2 * n);
    for rgsw in list {
        let (neg, pos) = rgsw_halves(rgsw);
        for half in [neg, pos] {
            let rows = gadget_rows(half);
            assert_eq!(rows.len(), levels, "unexpected gadget length");
            for row in rows {
                push_row_coeff::<F>(row, table, &mut out);
            }
        }
    }
    out
}

/// `DetectionKey` (key_gen/detection.rs:9-16) -> flat coefficient-form arrays.
pub fn flatten_detection_key(dk: &DetectionKey) -> FlatDetectionKey {
    let bsk1 = flatten_brk::<FirstLevelField>(dk.first_level_blind_rotation_key(), 4, sys::OMR_N1);
    assert_eq!(bsk1.len(), 512 * 8 * 2 * sys::OMR_N1);
    let bsk2 = flatten_brk::<SecondLevelField>(dk.second_level_blind_rotation_key(), 6, sys::OMR_N2);
    assert_eq!(bsk2.len(), sys::OMR_LWE2_N * 12 * 2 * sys::OMR_N2);

    let kk = dk.first_level_key_switching_key();
    let mut ksk = Vec::with_capacity(sys::OMR_N1 * 27 * (sys::OMR_LWE2_N + 1));
    for i in 0..sys::OMR_N1 {
        for j in 0..27 {
            let (a, b) = ksk_row(kk, i, j);
            assert_eq!(a.len(), sys::OMR_LWE2_N);
            ksk.extend_from_slice(a);
            ksk.push(b);
        }
    }

    let tk = dk.trace_key();
    let table2 = dk.second_level_blind_rotation_key().ntt_table();
    let mut trace = Vec::with_capacity(11 * 25 * 2 * sys::OMR_N2);
    let steps = trace_steps(tk);
    assert_eq!(steps.len(), 11, "log2(N2) automorphism keys");
    for g in steps {
        let rows = gadget_rows(g);
        assert_eq!(rows.len(), 25);
        for row in rows {
            push_row_coeff::<SecondLevelField>(row, table2, &mut trace);
        }
    }
    FlatDetectionKey { bsk1, ksk, bsk2, trace }
}

// ---- clues and ciphertexts ----------------------------------------------------------------------------------------------------

/// `CmLweCiphertext<u16>` x n -> (`a [n][512]`, `b [n][7]`).  `a()` / `b()` are the accessors `extract_all` is built on.
pub fn flatten_clues(clues: &[CmLweCiphertext<ClueValue>]) -> (Vec<u16>, Vec<u16>) {
    let mut a = Vec::with_capacity(clues.len() * sys::OMR_CLUE_N);
    let mut b = Vec::with_capacity(clues.len() * sys::OMR_CLUE_COUNT);
    for c in clues {
        assert_eq!(c.msg_count(), sys::OMR_CLUE_COUNT, "Invalid clue count."); // detector.rs:511
        assert_eq!(c.a().len(), sys::OMR_CLUE_N);
        a.extend_from_slice(c.a());
        b.extend_from_slice(c.b());
    }
    (a, b)
}

/// coefficient-form `[a | b]` (2 x 2048 u64, as the library returns it under `OMR_OUT_COEFF`) -> `NttRlwe<SecondLevelField>`
pub fn ntt_rlwe_from_coeff(words: &[u64], table: &<SecondLevelField as NttField>::Table) -> NttRlweCiphertext<SecondLevelField> {
    assert_eq!(words.len(), sys::OMR_PV_WORDS);
    let mut a = words[..sys::OMR_N2].to_vec();
    let mut b = words[sys::OMR_N2..].to_vec();
    table.transform_slice(a.as_mut_slice()); // the transform the reference packs with: detector.rs:325,435
    table.transform_slice(b.as_mut_slice());
    ntt_rlwe_from_parts::<SecondLevelField>(a, b)
}

/// `NttRlwe<SecondLevelField>` -> coefficient-form `[a | b]` appended to `out`
pub fn ntt_rlwe_to_coeff(ct: &NttRlweCiphertext<SecondLevelField>, table: &<SecondLevelField as NttField>::Table, out: &mut Vec<u64>) {
    push_row_coeff::<SecondLevelField>(ct, table, out);
}
