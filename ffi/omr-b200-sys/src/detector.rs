//! `GpuDetector`: `omr_core::Detector` (omr_core/src/detector.rs:35-453) over libomr_b200.so.
//!
//! Same method names, argument meaning and error behaviour (panics where the reference panics/asserts:
//! detector.rs:236, 511).  `&self` is shared between threads like the reference's `&Detector` (examples/omr.rs:160-164): the
//! context serialises callers internally, and the batch is the parallelism — prefer [`GpuDetector::detect_batch`] over
//! `par_iter().map(detect)`.

use std::{
    sync::Mutex,
    time::Duration,
};

use algebra::{ntt::NttTable, polynomial::FieldPolynomial, NttField};
use fhe_core::{CmLweCiphertext, NttRlweCiphertext};
use lattice::NttRlwe;
use omr_core::{ClueValue, DetectionKey, FirstLevelField, Payload, RetrievalParams, SecondLevelField, PAYLOAD_LENGTH};
use rand::{CryptoRng, Rng, SeedableRng};
use rand_distr::{Distribution, Uniform};

use crate::{check, flatten, last_error, sys};

/// `DetectTimeInfoPerMessage` (detector.rs:43-48), device time summed over the batch of the call.
#[derive(Debug, Clone, Copy, Default)]
pub struct DetectTimeInfoPerMessage {
    pub detect_time: Duration,
    pub first_level_bootstrapping_time: Duration,
    pub second_level_bootstrapping_time: Duration,
    pub trace_time: Duration,
}

/// Drop-in for `omr_core::Detector`, bound to one GPU.
pub struct GpuDetector {
    ctx: *mut sys::OmrCtx,
    detection_key: DetectionKey,
    first_level_lut: FieldPolynomial<FirstLevelField>,
    second_level_lut: FieldPolynomial<SecondLevelField>,
    /// `detect` = reset the store + detect one batch + read it back: one caller at a time
    store: Mutex<()>,
}

// SAFETY: the context is internally synchronised (include/omr_b200.h, "Concurrency"); the raw pointer is only handed to the library.
unsafe impl Send for GpuDetector {}
unsafe impl Sync for GpuDetector {}

impl GpuDetector {
    /// `Detector::new(detection_key)` (detector.rs:85-110) on GPU `OMR_B200_DEVICE` (default 0).
    /// Panics if the library cannot create a context — there is no CPU fallback.
    pub fn new(detection_key: DetectionKey) -> Self {
        let device = std::env::var("OMR_B200_DEVICE").ok().and_then(|v| v.parse().ok()).unwrap_or(0);
        Self::with_device(detection_key, device)
    }

    pub fn with_device(detection_key: DetectionKey, device: i32) -> Self {
        let flat = flatten::flatten_detection_key(&detection_key);
        let blobs = flat.blobs();
        let mut ctx = std::ptr::null_mut();
        // SAFETY: the four arrays outlive the call; the library copies them to the device
        let st = unsafe { sys::omr_ctx_create(device, &blobs, &mut ctx) };
        if st != sys::OMR_OK {
            panic!("omr_ctx_create failed (status {st}): {}", last_error(std::ptr::null()));
        }
        // SAFETY: ctx is valid from here on
        unsafe { assert_eq!(sys::omr_set_output_domain(ctx, sys::OMR_OUT_COEFF), sys::OMR_OK) };
        let mut l1 = vec![0u32; sys::OMR_N1];
        let mut l2 = vec![0u64; sys::OMR_N2];
        unsafe {
            assert_eq!(sys::omr_first_level_lut(ctx, l1.as_mut_ptr()), sys::OMR_OK);
            assert_eq!(sys::omr_second_level_lut(ctx, l2.as_mut_ptr()), sys::OMR_OK);
        }
        Self {
            ctx,
            detection_key,
            first_level_lut: FieldPolynomial::new(l1),
            second_level_lut: FieldPolynomial::new(l2),
            store: Mutex::new(()),
        }
    }

    /// detector.rs:112-114
    pub fn detect_key_size(&self) -> usize {
        unsafe { sys::omr_detect_key_size(self.ctx) }
    }
    /// detector.rs:118-120
    pub fn detection_key(&self) -> &DetectionKey {
        &self.detection_key
    }
    /// detector.rs:124-126 (built by the library exactly as detector.rs:457-476 + lut.rs:12-27 do)
    pub fn first_level_lut(&self) -> &FieldPolynomial<FirstLevelField> {
        &self.first_level_lut
    }
    /// detector.rs:130-132
    pub fn second_level_lut(&self) -> &FieldPolynomial<SecondLevelField> {
        &self.second_level_lut
    }

    fn table(&self) -> &<SecondLevelField as NttField>::Table {
        self.detection_key.second_level_blind_rotation_key().ntt_table()
    }

    fn check(&self, st: i32) {
        if let Err(e) = check(st, self.ctx) {
            panic!("{e}");
        }
    }

    /// `Detector::detect` (detector.rs:135-166) for one message.  Panics on a wrong clue count like the reference (:511).
    pub fn detect(&self, clues: &CmLweCiphertext<ClueValue>) -> NttRlweCiphertext<SecondLevelField> {
        self.detect_batch(std::slice::from_ref(clues)).pop().unwrap()
    }

    /// The batched form: replaces `clues_list.par_iter().map(|c| detector.detect(c)).collect()` (examples/omr.rs:160-164).
    /// The pertinency vector also stays resident on the GPU for [`Self::encode_pertinent_indices_resident`].
    pub fn detect_batch(&self, clues: &[CmLweCiphertext<ClueValue>]) -> Vec<NttRlweCiphertext<SecondLevelField>> {
        self.detect_batch_with_time_info(clues).0
    }

    /// `detect_with_time_info` (detector.rs:169-221)
    pub fn detect_with_time_info(&self, clues: &CmLweCiphertext<ClueValue>) -> (NttRlweCiphertext<SecondLevelField>, DetectTimeInfoPerMessage) {
        let (mut v, t) = self.detect_batch_with_time_info(std::slice::from_ref(clues));
        (v.pop().unwrap(), t)
    }

    pub fn detect_batch_with_time_info(&self, clues: &[CmLweCiphertext<ClueValue>]) -> (Vec<NttRlweCiphertext<SecondLevelField>>, DetectTimeInfoPerMessage) {
        let (a, b) = flatten::flatten_clues(clues);
        let mut pv = vec![0u64; clues.len() * sys::OMR_PV_WORDS];
        let mut t = sys::OmrStageTimes { detect_ms: 0.0, first_level_bootstrapping_ms: 0.0, second_level_bootstrapping_ms: 0.0, trace_ms: 0.0 };
        {
            let _g = self.store.lock().unwrap();
            // SAFETY: buffers are sized for `clues.len()` messages
            unsafe {
                self.check(sys::omr_pv_reset(self.ctx));
                self.check(sys::omr_detect_batch(self.ctx, a.as_ptr(), b.as_ptr(), clues.len(), 0, pv.as_mut_ptr(), &mut t));
            }
        }
        let table = self.table();
        let out = pv.chunks_exact(sys::OMR_PV_WORDS).map(|w| flatten::ntt_rlwe_from_coeff(w, table)).collect();
        let ms = |x: f32| Duration::from_secs_f64(x as f64 * 1e-3);
        (out, DetectTimeInfoPerMessage {
            detect_time: ms(t.detect_ms),
            first_level_bootstrapping_time: ms(t.first_level_bootstrapping_ms),
            second_level_bootstrapping_time: ms(t.second_level_bootstrapping_ms),
            trace_time: ms(t.trace_ms),
        })
    }

    fn c_params(rp: &RetrievalParams<SecondLevelField>) -> sys::OmrRetrievalParams {
        let mut c = sys::OmrRetrievalParams {
            index_modulus: 0, polynomial_size: 0, bucket_count_per_segment: 0, slots_per_bucket: 0, slots_per_segment: 0, segment_count: 0,
            segment_per_cipher: 0, max_encode_indices_cipher_count: 0, pertinent_count: 0, combination_count: 0, cmb_count_per_cipher: 0,
            all_payloads_count: 0,
        };
        // RetrievalParams::new(257, 2048, D, k, 130, 25, 2) (secret.rs:189-209) recomputed by the library, then checked against `rp`
        unsafe { assert_eq!(sys::omr_retrieval_params_init(rp.all_payloads_count() as u64, rp.pertinent_count() as u32, &mut c), sys::OMR_OK) };
        assert_eq!(c.polynomial_size as usize, rp.polynomial_size());
        assert_eq!(c.slots_per_bucket as usize, rp.slots_per_bucket());
        assert_eq!(c.slots_per_segment as usize, rp.slots_per_segment());
        assert_eq!(c.max_encode_indices_cipher_count as usize, rp.max_encode_indices_cipher_count());
        assert_eq!(c.combination_count as usize, rp.combination_count());
        assert_eq!(c.cmb_count_per_cipher as usize, rp.cmb_count_per_cipher());
        c
    }

    /// load a pertinency vector the caller holds into the resident store (coefficient form through Primus-fhe's own table)
    fn load_store(&self, pertinency_vector: &[NttRlweCiphertext<SecondLevelField>]) {
        let table = self.table();
        let mut flat = Vec::with_capacity(pertinency_vector.len() * sys::OMR_PV_WORDS);
        for ct in pertinency_vector {
            flatten::ntt_rlwe_to_coeff(ct, table, &mut flat);
        }
        unsafe { self.check(sys::omr_pv_load(self.ctx, flat.as_ptr(), pertinency_vector.len(), 0)) };
    }

    /// `Detector::encode_pertinent_indices` (detector.rs:223-339): one index-digest ciphertext; the reference calls it
    /// `max_encode_indices_cipher_count` times (examples/omr.rs:180-183), each call with fresh random buckets (:262).
    pub fn encode_pertinent_indices(&self, retrieval_params: RetrievalParams<SecondLevelField>, pertinency_vector: &[NttRlweCiphertext<SecondLevelField>]) -> NttRlwe<SecondLevelField> {
        assert_eq!(retrieval_params.polynomial_size(), self.table().dimension()); // detector.rs:236
        let _g = self.store.lock().unwrap();
        self.load_store(pertinency_vector);
        self.encode_indices_locked(&retrieval_params, 1).pop().unwrap()
    }

    /// All index ciphertexts at once over the vector left resident by the last `detect_batch` (no 32 KiB-per-message upload).
    pub fn encode_pertinent_indices_resident(&self, retrieval_params: RetrievalParams<SecondLevelField>) -> Vec<NttRlwe<SecondLevelField>> {
        let _g = self.store.lock().unwrap();
        self.encode_indices_locked(&retrieval_params, retrieval_params.max_encode_indices_cipher_count())
    }

    fn encode_indices_locked(&self, rp: &RetrievalParams<SecondLevelField>, n: usize) -> Vec<NttRlwe<SecondLevelField>> {
        let c = Self::c_params(rp);
        let seed: u64 = rand::thread_rng().gen(); // bucket choice: thread_rng in the reference (detector.rs:262)
        let mut out = vec![0u64; n * sys::OMR_PV_WORDS];
        unsafe { self.check(sys::omr_encode_indices(self.ctx, &c, seed, 0, n as u32, out.as_mut_ptr())) };
        let table = self.table();
        out.chunks_exact(sys::OMR_PV_WORDS).map(|w| flatten::ntt_rlwe_from_coeff(w, table)).collect()
    }

    /// `Detector::encode_pertinent_payloads` (detector.rs:341-453).  The weights are drawn here exactly as the reference does
    /// (:376-387: `Uniform::new(0, p).sample_iter(rng)`, row-major `[combination][message]`), so `Retriever::decode_digest`
    /// regenerates the same matrix from the same seed (retriever.rs:215-226).
    pub fn encode_pertinent_payloads<R>(&self, pertinency_vector: &[NttRlweCiphertext<SecondLevelField>], payloads: &[Payload], combination_count: usize,
                                        cmb_count_per_cipher: usize, rng: &mut R) -> Vec<NttRlweCiphertext<SecondLevelField>>
    where
        R: Rng + SeedableRng + CryptoRng,
    {
        assert_eq!(pertinency_vector.len(), payloads.len());
        let count = payloads.len();
        let p = self.detection_key.params().output_plain_modulus_value() as u16;
        let weights: Vec<u16> = Uniform::new(0u16, p).sample_iter(&mut *rng).take(combination_count * count).collect();
        let mut flat_payloads = Vec::with_capacity(count * PAYLOAD_LENGTH);
        for pl in payloads {
            flat_payloads.extend_from_slice(&pl.0);
        }
        let n_cipher = combination_count.div_ceil(cmb_count_per_cipher);
        let mut out = vec![0u64; n_cipher * sys::OMR_PV_WORDS];
        {
            let _g = self.store.lock().unwrap();
            self.load_store(pertinency_vector);
            unsafe {
                self.check(sys::omr_encode_payloads(self.ctx, flat_payloads.as_ptr(), count, weights.as_ptr(), combination_count, count,
                                                    n_cipher as u32, cmb_count_per_cipher as u32, out.as_mut_ptr()));
            }
        }
        let table = self.table();
        out.chunks_exact(sys::OMR_PV_WORDS).map(|w| flatten::ntt_rlwe_from_coeff(w, table)).collect()
    }
}

impl Drop for GpuDetector {
    fn drop(&mut self) {
        // SAFETY: the context is not used after this
        unsafe { sys::omr_ctx_destroy(self.ctx) }
    }
}
