"""Constants of the one parameter set the reference defines (omr_core/src/parameters/mod.rs:39-105) and the digest
layout (omr_core/src/parameters/retrieval_params.rs:50-106)."""
from dataclasses import dataclass

PAYLOAD_LENGTH = 612                      # payload.rs:8


@dataclass(frozen=True)
class OmrParameters:
    """OmrParameters::new() — parameters/mod.rs:39-105."""
    clue_dimension: int = 512
    clue_cipher_modulus: int = 2048
    clue_plain_modulus: int = 8
    clue_count: int = 7
    first_level_ring_dimension: int = 1024
    first_level_modulus: int = 134215681
    first_level_basis: tuple = (5, 4)          # (log_basis, levels)
    key_switching_basis: tuple = (1, 27)
    intermediate_lwe_dimension: int = 670
    intermediate_lwe_cipher_modulus: int = 4096
    intermediate_lwe_plain_modulus: int = 32
    second_level_ring_dimension: int = 2048
    second_level_modulus: int = 1125899906826241
    second_level_basis: tuple = (7, 6)
    trace_basis: tuple = (2, 25)
    output_plain_modulus_value: int = 257


class RetrievalParams:
    """RetrievalParams::new — retrieval_params.rs:50-106; the defaults are those of
    SecretKeyPack::generate_retriever (key_gen/secret.rs:189-209)."""

    def __init__(self, all_payloads_count, pertinent_count, index_modulus=257, polynomial_size=2048,
                 bucket_count_per_segment=130, segment_count=25, cmb_count_per_cipher=2):
        if index_modulus != 257 or polynomial_size != 2048:
            raise ValueError("only index_modulus 257 / polynomial_size 2048 (the reference's parameter set)")
        self.index_modulus = index_modulus
        self.polynomial_size = polynomial_size
        self.bucket_count_per_segment = bucket_count_per_segment
        self.segment_count = segment_count
        self.cmb_count_per_cipher = cmb_count_per_cipher
        self.all_payloads_count = all_payloads_count
        self.pertinent_count = pertinent_count
        e, pw = 1, index_modulus
        while pw < all_payloads_count:
            pw *= index_modulus
            e += 1
        self.slots_per_bucket = e + 1
        self.slots_per_segment = self.slots_per_bucket * bucket_count_per_segment
        self.segment_per_cipher = polynomial_size // self.slots_per_segment
        self.max_encode_indices_cipher_count = segment_count // self.segment_per_cipher
        self.combination_count = pertinent_count + 5          # index_modulus is not a power of two

    @property
    def payload_cipher_count(self):
        return -(-self.combination_count // self.cmb_count_per_cipher)

    def to_c(self):
        from ._lib import RetrievalParamsC
        c = RetrievalParamsC()
        for name, _ in RetrievalParamsC._fields_:
            setattr(c, name, getattr(self, name))
        return c
