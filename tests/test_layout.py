"""Python model of the NTT pass geometry and padded exchange layouts of csrc/ntt.cuh (GeoL1 / GeoL2): the model's
configuration is read from the typedefs in the header, every pass is a partition of the N coefficients, per-thread
shared-memory offsets are thread-independent constants, every warp-wide access is bank-conflict free, and the multi-pass
register schedule computes the oracle's Cooley-Tukey transform."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import layout_check as LC  # noqa: E402

FIELDS = {"L1": (134215681, 4073518), "L2": (1125899906826241, 765727830662934)}


def _header_config(name):
    src = open(os.path.join(ROOT, "tfhe-omr_b200", "csrc", "ntt.cuh")).read()
    m = re.search(r"typedef GeoT<([^>]*)>\s+Geo%s;" % name, src)
    toks = [x.strip() for x in m.group(1).split(",")]
    blocked = toks[-1] == "true"
    v = [int(x) for x in toks if x not in ("true", "false")]
    N, logn, NT, E, npass = v[:5]
    ns = v[5:5 + npass]
    pads = [(v[9 + 2 * x], v[10 + 2 * x]) for x in range(npass - 1)]
    assert N == 1 << logn and NT * E == N and sum(ns) == logn
    return N, NT, E, ns, pads, blocked


@pytest.mark.parametrize("name", ["L1", "L2"])
def test_model_matches_header(name):
    N, NT, E, ns, pads, blocked = _header_config(name)
    cfg = LC.CONFIGS[name]
    assert (cfg[0], cfg[1], cfg[2], cfg[3], cfg[5], cfg[6]) == (N, NT, E, ns, pads, blocked)


@pytest.mark.parametrize("name", ["L1", "L2"])
def test_pass_partition_offsets_and_bank_conflicts(name, capsys):
    assert LC.check(name)                                  # asserts partitions / constant offsets, returns conflict-freedom


@pytest.mark.parametrize("name", ["L1", "L2"])
def test_register_schedule_matches_reference_ntt(name):
    """the multi-pass register schedule (Pass<GEO,P>::idx, twiddle index (1 << (s0+l)) + (j << l) + sb) computes the same
    in-place Cooley-Tukey transform as the oracle's loop nest (natural in, bit-reversed out)"""
    cfg = LC.CONFIGS[name]
    N, NT, E, ns = cfg[:4]
    logn = N.bit_length() - 1
    q, psi = FIELDS[name]
    rng = np.random.default_rng(0)
    a = [int(v) for v in rng.integers(0, 1 << 26, N)]
    brv = lambda x, b: int(format(x, f"0{b}b")[::-1], 2)
    tw = [pow(psi, brv(i, logn), q) for i in range(N)]
    ref = a[:]
    t, m = N, 1
    while m < N:
        t >>= 1
        for i in range(m):
            for j in range(2 * i * t, 2 * i * t + t):
                u, v = ref[j], ref[j + t] * tw[m + i] % q
                ref[j], ref[j + t] = (u + v) % q, (u - v) % q
        m <<= 1
    x = a[:]
    for p in range(len(ns)):
        s0, EP = sum(ns[:p]), 1 << ns[p]
        stride = (N >> s0) // EP
        for tt in range(NT):
            for g in range(E // EP):
                j = LC.vthread(cfg, p, tt, g) // stride
                pos = [LC.idx(cfg, p, tt, g * EP + kk) for kk in range(EP)]
                for l in range(ns[p]):
                    half = EP >> (l + 1)
                    for sb in range(1 << l):
                        w = tw[(1 << (s0 + l)) + (j << l) + sb]
                        for h in range(half):
                            lo, hi = pos[sb * 2 * half + h], pos[sb * 2 * half + h + half]
                            u, v = x[lo], x[hi] * w % q
                            x[lo], x[hi] = (u + v) % q, (u - v) % q
    assert x == ref


def test_no_exchange_hazards_under_arbitrary_warp_interleavings():
    """the warp-blocked level-2 geometry replaces two of three exchange barriers by __syncwarp(): replay the store/sync/load
    sequences of the kernels under random warp schedules and require every load to see its own store (scripts/layout_check.py)"""
    progs = {"K3 step": ["F2"] * 6 + ["I2", "B"] + ["F2"] * 2 + ["I2", "B"], "K4 step + final": ["F2"] * 3 + ["I2", "B", "B", "F2"],
             "pack": ["F"] * 4, "cluster": ["F2", "I", "B", "F2", "I", "B"], "ntt_kernel": ["I", "F"], "mixed": ["F", "I", "F2", "I2", "F", "I"]}
    for name, prog in progs.items():
        assert LC.race_check("L2", prog, schedules=25) is None, name
    assert LC.race_check("L1", ["F", "I", "I2", "F"], schedules=5) is None
    # K1 step: four digit pairs through the two buffers, the CTA barrier, both inverse transforms, the group barrier
    assert LC.race_check("L1", (["F2"] * 4 + ["B", "I2", "B"]) * 2, schedules=10) is None, "K1 step"
    # the detector itself: the unpadded first exchange lets warp w's private region overlap warp w+1's and must be flagged
    saved = LC.CONFIGS["L2"]
    try:
        LC.CONFIGS["L2"] = saved[:5] + ([(0, 0)] + saved[5][1:],) + saved[6:]
        assert LC.race_check("L2", ["F2", "I2", "B"], schedules=25) is not None
    finally:
        LC.CONFIGS["L2"] = saved
