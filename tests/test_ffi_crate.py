"""ffi/omr-b200-sys cannot be compiled here (no Rust toolchain), so this test keeps it honest against the C ABI:
* src/sys.rs is exactly what scripts/gen_rust_sys.py makes from include/omr_b200.h;
* an independent parse of the `extern "C"` block agrees with the header, symbol by symbol and type by type;
* the repr(C) structs have the sizes and field offsets gcc gives the C structs;
* every `sys::` item the hand-written modules use exists, and GpuDetector carries the reference's method signatures."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CRATE = os.path.join(ROOT, "ffi", "omr-b200-sys")
HDR = os.path.join(ROOT, "include", "omr_b200.h")

C2R = {"int": "c_int", "size_t": "usize", "uint8_t": "u8", "uint16_t": "u16", "uint32_t": "u32", "uint64_t": "u64", "int32_t": "i32",
       "double": "f64", "float": "f32", "char": "c_char", "void": "c_void", "omr_ctx": "OmrCtx", "omr_key_blobs": "OmrKeyBlobs",
       "omr_stage_times": "OmrStageTimes", "omr_retrieval_params": "OmrRetrievalParams", "omr_blob_header": "OmrBlobHeader",
       "omr_secret_key": "OmrSecretKey"}


def _c_decls():
    h = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    out = {}
    for ret, name, args in re.findall(r"^\s*((?:const\s+)?[A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)(omr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", h, flags=re.M):
        params = [re.sub(r"\s*[A-Za-z_][A-Za-z0-9_]*$", "", " ".join(a.split())) for a in args.split(",") if a.strip() and a.strip() != "void"]
        out[name] = (" ".join(ret.split()), params)
    return out


def _norm_c(t):
    """C type -> canonical Rust spelling, written independently of the generator: read the declarator right to left"""
    toks = t.replace("*", " * ").split()
    base = [x for x in toks if x not in ("const", "*")][0]
    r = C2R[base]
    # qualifiers: `const` before the first * binds to the base; a `const` after a * binds to that pointer
    first_star = toks.index("*") if "*" in toks else len(toks)
    const_here = "const" in toks[:first_star]
    i = first_star
    while i < len(toks):
        r = ("*const " if const_here else "*mut ") + r
        const_here = i + 1 < len(toks) and toks[i + 1] == "const"
        i += 2 if const_here else 1
    return r


def _rust_decls():
    src = open(os.path.join(CRATE, "src", "sys.rs")).read()
    block = src[src.index('extern "C" {'):]
    out = {}
    for name, args, ret in re.findall(r"pub fn (omr_[a-z0-9_]+)\((.*?)\)(?: -> ([^;]+))?;", block):
        params = [a.split(":", 1)[1].strip() for a in args.split(", ") if a.strip()]
        out[name] = (ret.strip() if ret else "()", params)
    return out


def test_sys_rs_is_generated_from_the_header():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_rust_sys.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_extern_block_matches_header_symbol_by_symbol():
    c, r = _c_decls(), _rust_decls()
    assert set(c) == set(r), set(c) ^ set(r)
    from tfhe_omr_b200 import _lib
    assert set(c) == set(_lib.EXPORTS)                                   # and the Python binding lists the same symbols
    for name, (cret, cparams) in c.items():
        rret, rparams = r[name]
        assert (rret == "()") == (cret == "void"), name
        if cret != "void":
            assert rret == _norm_c(cret), (name, cret, rret)
        assert [_norm_c(p) for p in cparams] == rparams, (name, cparams, rparams)


def test_repr_c_structs_have_the_c_layout(tmp_path):
    """sizeof / offsetof from gcc vs the layout repr(C) gives the Rust fields (natural alignment)"""
    src = open(os.path.join(CRATE, "src", "sys.rs")).read()
    size = {"u8": 1, "u16": 2, "u32": 4, "i32": 4, "f32": 4, "u64": 8, "f64": 8, "usize": 8}
    cname = {v: k for k, v in C2R.items() if k.startswith("omr_")}
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HDR}"', "int main(void) {"]
    want = {}
    for rname, body in re.findall(r"pub struct (Omr[A-Za-z]+) \{\n(.*?)\n\}", src, flags=re.S):
        if rname == "OmrCtx":
            continue
        off, align, fields = 0, 1, []
        for f, t in re.findall(r"pub ([a-z0-9_]+): ([^,]+),", body):
            m = re.match(r"\[(\w+); (\d+)\]", t)
            sz, al = (size[m.group(1)] * int(m.group(2)), size[m.group(1)]) if m else ((8, 8) if t.startswith("*") else (size[t], size[t]))
            off = -(-off // al) * al
            fields.append((f, off)); off += sz; align = max(align, al)
        want[rname] = (-(-off // align) * align, fields)
        prog.append(f'printf("{rname} %zu", sizeof({cname[rname]}));')
        for f, _ in fields:
            prog.append(f'printf(" %zu", offsetof({cname[rname]}, {f}));')
        prog.append('printf("\\n");')
    prog += ["return 0; }"]
    c = tmp_path / "layout.c"; c.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-o", str(exe), str(c)])
    got = {}
    for line in subprocess.check_output([str(exe)], text=True).splitlines():
        p = line.split(); got[p[0]] = [int(x) for x in p[1:]]
    assert set(got) == set(want) and len(want) == 5
    for rname, (sz, fields) in want.items():
        assert got[rname] == [sz] + [o for _, o in fields], rname


def test_handwritten_modules_only_use_existing_symbols_and_mirror_the_reference_signatures():
    sysrs = open(os.path.join(CRATE, "src", "sys.rs")).read()
    items = set(re.findall(r"pub (?:fn|const|struct) ([A-Za-z0-9_]+)", sysrs))
    used = set()
    for f in ("lib.rs", "blob.rs", "flatten.rs", "detector.rs"):
        used |= set(re.findall(r"(?<![A-Za-z0-9_])sys::([A-Za-z0-9_]+)", open(os.path.join(CRATE, "src", f)).read()))
    for f in os.listdir(os.path.join(CRATE, "examples")):
        used |= set(re.findall(r"(?<![A-Za-z0-9_])sys::([A-Za-z0-9_]+)", open(os.path.join(CRATE, "examples", f)).read()))
    used -= {"OmrCtx"} - items
    assert used <= items, used - items
    det = " ".join(open(os.path.join(CRATE, "src", "detector.rs")).read().split())
    for sig in (                                                         # omr_core/src/detector.rs:85, 112-132, 135-138, 169-175, 223-227, 341-351
        "pub fn new(detection_key: DetectionKey) -> Self",
        "pub fn detect_key_size(&self) -> usize",
        "pub fn detection_key(&self) -> &DetectionKey",
        "pub fn first_level_lut(&self) -> &FieldPolynomial<FirstLevelField>",
        "pub fn second_level_lut(&self) -> &FieldPolynomial<SecondLevelField>",
        "pub fn detect(&self, clues: &CmLweCiphertext<ClueValue>) -> NttRlweCiphertext<SecondLevelField>",
        "pub fn detect_with_time_info(&self, clues: &CmLweCiphertext<ClueValue>) -> (NttRlweCiphertext<SecondLevelField>, DetectTimeInfoPerMessage)",
        "pub fn encode_pertinent_indices(&self, retrieval_params: RetrievalParams<SecondLevelField>, pertinency_vector: &[NttRlweCiphertext<SecondLevelField>]) -> NttRlwe<SecondLevelField>",
        "pub fn encode_pertinent_payloads<R>(&self, pertinency_vector: &[NttRlweCiphertext<SecondLevelField>], payloads: &[Payload], combination_count: usize, cmb_count_per_cipher: usize, rng: &mut R) -> Vec<NttRlweCiphertext<SecondLevelField>>",
    ):
        assert sig in det, sig
    # the blob kinds the dumper writes are the files the vector test reads
    dump = open(os.path.join(CRATE, "examples", "dump_vectors.rs")).read()
    import test_ref_vectors as T
    for f in T.FILES:
        assert f'"{f}.omrb"' in dump, f
