"""Builds libomr_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "lib", "libomr_b200.so")
SOURCES = [os.path.join(HERE, "csrc", f) for f in ("capi.cu", "ks_gemm.cu", "blob.cu", "nccl_dl.cu", "nccl_dl.hpp", "kernels.cuh", "ntt.cuh", "field.cuh")] + \
          [os.path.join(ROOT, "include", "omr_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-cudart", "static"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libomr_b200.so cannot be built (there is no CPU fallback)")


def _cutlass_include():
    """CUTLASS / CuTe header trees vendored in the image (used by csrc/ks_gemm.cu only).  None -> that unit is compiled
    without the tensor-core GEMM and the library keeps its CUDA-core key switch."""
    cands = [os.environ.get("OMR_CUTLASS_DIR", "")]
    try:
        import importlib.util
        for pkg, sub in (("flashinfer", "data/cutlass"), ("tilelang", "3rdparty/cutlass")):
            spec = importlib.util.find_spec(pkg)
            if spec and spec.submodule_search_locations:
                cands.append(os.path.join(list(spec.submodule_search_locations)[0], sub))
    except Exception:
        pass
    for c in cands:
        if c and os.path.exists(os.path.join(c, "include", "cutlass", "gemm", "collective", "collective_builder.hpp")) and \
                os.path.exists(os.path.join(c, "tools", "util", "include", "cutlass", "util", "packed_stride.hpp")):
            return c
    return None


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in SOURCES)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cutlass = _cutlass_include()
    extra = ["-DOMR_HAVE_CUTLASS", "--expt-relaxed-constexpr", "-diag-suppress", "20012", "-I" + os.path.join(cutlass, "include"),
             "-I" + os.path.join(cutlass, "tools", "util", "include")] if cutlass else []
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB] + [s for s in SOURCES if s.endswith(".cu")] + ["-ldl"]
    subprocess.check_call(cmd, cwd=HERE)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
