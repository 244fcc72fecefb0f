"""Python model of the NTT pass geometry and exchange-buffer swizzles of csrc/ntt.cuh: every pass is a partition of
the N coefficients and every warp-wide shared-memory access is bank-conflict free."""
import numpy as np
import pytest


def swz1(x):
    return x ^ ((((x >> 5) & 1) * 3) | (((x >> 6) & 1) * 4) | (((x >> 7) & 1) * 24))


def swz2(x):
    return x ^ (((x >> 4) & 1) | (((x >> 5) & 1) * 6) | (((x >> 6) & 1) * 8))


def idx(N, logn, p, t, k):
    NT = N // 8
    ns = 3 if p < 3 else logn - 9
    s0 = 3 * p
    EP = 1 << ns
    blk = N >> s0
    stride = blk // EP
    g, kk = divmod(k, EP)
    vt = t + NT * g
    j, i = divmod(vt, stride)
    return j * blk + i + kk * stride


@pytest.mark.parametrize("N,logn,swz,lanes,width", [(1024, 10, swz1, 32, 4), (2048, 11, swz2, 16, 8)])
def test_pass_partition_and_bank_conflicts(N, logn, swz, lanes, width):
    NT = N // 8
    assert sorted(swz(x) for x in range(N)) == list(range(N))           # the swizzle is a permutation
    for p in range(4):
        seen = sorted(idx(N, logn, p, t, k) for t in range(NT) for k in range(8))
        assert seen == list(range(N))
        for k in range(8):
            for w0 in range(0, NT, lanes):                                 # one (half-)warp per shared-memory transaction
                addrs = [swz(idx(N, logn, p, t, k)) * width for t in range(w0, w0 + lanes)]
                banks = [(a // width) % (128 // width) for a in addrs]
                assert len(set(banks)) == lanes, (p, k, w0)


def test_bitrev_butterfly_schedule_matches_reference_ntt():
    """the 4-pass register schedule computes the same in-place Cooley-Tukey transform as the oracle's loop nest"""
    N, logn, q = 1024, 10, 134215681
    psi = 4073518
    rng = np.random.default_rng(0)
    a = [int(v) for v in rng.integers(0, q, N)]
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
    NT = N // 8
    for p in range(4):
        ns = 3 if p < 3 else logn - 9
        s0, EP = 3 * p, 1 << ns
        stride = (N >> s0) // EP
        for tt in range(NT):
            for g in range(8 // EP):
                j = (tt + NT * g) // stride
                pos = [idx(N, logn, p, tt, g * EP + kk) for kk in range(EP)]
                for l in range(ns):
                    half = EP >> (l + 1)
                    for sb in range(1 << l):
                        w = tw[(1 << (s0 + l)) + (j << l) + sb]
                        for h in range(half):
                            lo, hi = pos[sb * 2 * half + h], pos[sb * 2 * half + h + half]
                            u, v = x[lo], x[hi] * w % q
                            x[lo], x[hi] = (u + v) % q, (u - v) % q
    assert x == ref
