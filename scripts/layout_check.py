"""Model of the padded exchange layouts of csrc/ntt.cuh: for each configuration verifies that every pass is a partition of
the N coefficients, that physical offsets are (per-thread base + thread-independent constant), and that every warp-wide
shared-memory access is bank-conflict free (4-byte elements: 32 lanes/phase; 8-byte: 16; 16-byte vectors: 8)."""
import sys

CONFIGS = {
    # name: (N, NT, E, stages per pass, element bytes, [(A, S) padding of exchange x: phys = idx + A*(idx >> S)])
    "L1": (1024, 64, 16, [4, 4, 2], 4, [(4, 6), (4, 6)]),
    "L2": (2048, 256, 8, [3, 3, 3, 2], 8, [(0, 0), (4, 5), (2, 4)]),
}


def idx(cfg, p, t, k):
    N, NT, E, ns, _, _ = cfg
    s0 = sum(ns[:p]); EP = 1 << ns[p]; blk = N >> s0; stride = blk // EP
    g, kk = divmod(k, EP)
    vt = t + NT * g
    j, i = divmod(vt, stride)
    return j * blk + i + kk * stride


def phys(cfg, x, i):
    A, S = cfg[5][x]
    return i + A * (i >> S) if A else i


def check(name):
    cfg = CONFIGS[name]
    N, NT, E, ns, eb, pads = cfg
    ok = True
    for p in range(len(ns)):
        assert sorted(idx(cfg, p, t, k) for t in range(NT) for k in range(E)) == list(range(N)), (name, p)
        EP = 1 << ns[p]; stride = (N >> sum(ns[:p])) // EP
        for x, role in ((p - 1, "read"), (p, "write")):
            if x < 0 or x >= len(pads):
                continue
            offs = [phys(cfg, x, idx(cfg, p, 0, k)) - phys(cfg, x, idx(cfg, p, 0, 0)) for k in range(E)]
            for t in range(NT):
                b = phys(cfg, x, idx(cfg, p, t, 0))
                assert [phys(cfg, x, idx(cfg, p, t, k)) - b for k in range(E)] == offs, (name, p, role, t)
            vec = EP if stride == 1 else 1                     # contiguous group -> vector access (<= 16 bytes each)
            vbytes = min(16, vec * eb); per = vbytes // eb
            lanes = 128 // vbytes if vbytes > 4 else 32
            worst = 1
            for k in range(0, E, per):
                for w0 in range(0, NT, lanes):
                    addrs = [phys(cfg, x, idx(cfg, p, t, k)) * eb for t in range(w0, min(NT, w0 + lanes))]
                    assert all(a % vbytes == 0 for a in addrs)
                    units = [(a // vbytes) % (128 // vbytes) for a in addrs]
                    worst = max(worst, max(units.count(u) for u in set(units)))
            print(f"{name} pass {p} {role:5s} exchange {x}: vector {vbytes:2d} B, {lanes} lanes/phase, worst conflict degree {worst}")
            ok &= worst == 1
        assert len({phys(cfg, x, i) for i in range(N)}) == N if (x := min(p, len(pads) - 1)) >= 0 else True
    size = [max(phys(cfg, x, i) for i in range(N)) + 1 for x in range(len(pads))]
    print(f"{name}: buffer elements per exchange {size}")
    return ok


if __name__ == "__main__":
    good = all(check(n) for n in CONFIGS)
    sys.exit(0 if good else 1)
