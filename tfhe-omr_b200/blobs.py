"""Versioned flat blobs for the data either side of the hot path.

The reference serialises nothing (it only reports byte counts through `Size`: key_gen/detection.rs:81-88, sender.rs:36);
these little-endian containers let the Rust shim, the oracle and the GPU library exchange keys, clues, pertinency vectors
and digests, and make runs replayable.  Layouts are exactly the arrays of include/omr_b200.h.

    header (64 bytes): magic b"OMRB200\\0" | u32 version | u32 kind | u64 count | u64 index0 | u64 aux | u64 payload bytes |
                       u32 domain (0 = this library's NTT ordering, 1 = coefficient form) | 12 B reserved
    payload: the arrays of the kind, in the order listed in KINDS, C-contiguous, little-endian

The same container is read and written by libomr_b200.so (csrc/blob.cu: omr_blob_read / omr_blob_write /
omr_ctx_create_from_blob) and by the CPU restatement used as test infrastructure; tests/test_blobs.py cross-checks the three.
"""
import struct

import numpy as np

MAGIC = b"OMRB200\0"
VERSION = 1
_HDR = struct.Struct("<8sIIQQQQI12x")
DOMAIN_NTT_NATIVE, DOMAIN_COEFF = 0, 1

# kind -> (name, [(field, dtype, shape with -1 = count)])
KINDS = {
    1: ("detection_key", [("bsk1", np.uint32, (512, 8, 2, 1024)), ("ksk", np.uint32, (1024, 27, 671)),
                          ("bsk2", np.uint64, (670, 12, 2, 2048)), ("trace", np.uint64, (11, 25, 2, 2048))]),   # aux: key flags (0 NTT, 1 coeff)
    2: ("clues", [("a", np.uint16, (-1, 512)), ("b", np.uint16, (-1, 7))]),                                     # CmLweCiphertext<u16> x count
    3: ("pertinency_vector", [("pv", np.uint64, (-1, 2, 2048))]),                                              # NttRlweCiphertext<F2> x count
    4: ("digest", [("ct", np.uint64, (-1, 2, 2048))]),                                                         # aux: number of index ciphertexts
    5: ("payloads", [("payloads", np.uint16, (-1, 612))]),
    6: ("secret_key", [("s0", np.int32, (512,)), ("z1", np.int32, (1024,)), ("s2", np.int32, (670,)), ("z2", np.int32, (2048,))]),   # test vectors only
    7: ("rlwe1", [("ct", np.uint32, (-1, 2, 1024))]),                                                          # sum of the 7 L1 accumulators (detector.rs:556)
    8: ("lwe2", [("ct", np.uint32, (-1, 671))]),                                                               # after KS + mod switch + offset (detector.rs:560-596)
    9: ("rlwe2", [("ct", np.uint64, (-1, 2, 2048))]),                                                          # after the L2 blind rotation (detector.rs:623)
    10: ("clue_key", [("pa", np.uint16, (512,)), ("pb", np.uint16, (512,))]),
}
_BY_NAME = {v[0]: k for k, v in KINDS.items()}


def dump(path, kind, arrays, count=0, index0=0, aux=0, domain=DOMAIN_NTT_NATIVE):
    """Write one blob.  `arrays` maps field name -> array."""
    kid = _BY_NAME[kind]
    fields = KINDS[kid][1]
    parts = []
    for name, dt, shape in fields:
        arr = np.ascontiguousarray(arrays[name], dtype=np.dtype(dt).newbyteorder("<"))
        want = tuple(count if s == -1 else s for s in shape)
        if arr.shape != want:
            raise ValueError(f"{kind}.{name}: shape {arr.shape}, expected {want}")
        parts.append(arr)
    nbytes = sum(p.nbytes for p in parts)
    with open(path, "wb") as f:
        f.write(_HDR.pack(MAGIC, VERSION, kid, count, index0, aux, nbytes, domain))
        for p in parts:
            f.write(p.tobytes())


def load(path, mmap=False):
    """Read one blob -> (kind name, dict of arrays, header dict)."""
    with open(path, "rb") as f:
        hdr = f.read(_HDR.size)
        if len(hdr) != _HDR.size:
            raise ValueError("truncated header")
        magic, version, kid, count, index0, aux, nbytes, domain = _HDR.unpack(hdr)
        if magic != MAGIC:
            raise ValueError("not an OMRB200 blob")
        if version != VERSION:
            raise ValueError(f"unsupported blob version {version}")
        if kid not in KINDS:
            raise ValueError(f"unknown blob kind {kid}")
        name, fields = KINDS[kid]
        out, off = {}, _HDR.size
        for fname, dt, shape in fields:
            shp = tuple(count if s == -1 else s for s in shape)
            n = int(np.prod(shp)) * np.dtype(dt).itemsize
            if mmap:
                out[fname] = np.memmap(path, dtype=np.dtype(dt).newbyteorder("<"), mode="r", offset=off, shape=shp)
            else:
                f.seek(off)
                buf = f.read(n)
                if len(buf) != n:
                    raise ValueError("truncated payload")
                out[fname] = np.frombuffer(buf, dtype=np.dtype(dt).newbyteorder("<")).reshape(shp)
            off += n
        if off - _HDR.size != nbytes:
            raise ValueError("payload size mismatch")
        f.seek(0, 2)
        if f.tell() != off:
            raise ValueError("trailing bytes after the payload")
    return name, out, {"version": version, "count": count, "index0": index0, "aux": aux, "domain": domain}
