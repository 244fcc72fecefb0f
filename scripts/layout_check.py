"""Model of the padded exchange layouts of csrc/ntt.cuh: for each configuration verifies that every pass is a partition of
the N coefficients, that physical offsets are (per-thread base + thread-independent constant), and that every warp-wide
shared-memory access is bank-conflict free (4-byte elements: 32 lanes/phase; 8-byte: 16; 16-byte vectors: 8)."""
import sys

CONFIGS = {
    # name: (N, NT, E, stages per pass, element bytes, [(A, S) padding of exchange x: phys = idx + A*(idx >> S)], warp-blocked)
    "L1": (1024, 64, 16, [4, 4, 2], 4, [(4, 6), (4, 6)], False),
    "L2": (2048, 256, 8, [3, 3, 3, 2], 8, [(32, 8), (4, 5), (2, 4)], True),
}


def vthread(cfg, p, t, g):
    """virtual thread of register group g in pass p: NT-strided, or — warp-blocked geometry, last pass — the g-th 32-thread slice
    of the warp's own block (ntt.cuh: Pass::vt)"""
    N, NT, E, ns = cfg[:4]
    if cfg[6] and p == len(ns) - 1:
        G = E >> ns[p]
        return (t // 32) * 32 * G + (t % 32) + 32 * g
    return t + NT * g


def idx(cfg, p, t, k):
    N, NT, E, ns = cfg[:4]
    s0 = sum(ns[:p]); EP = 1 << ns[p]; blk = N >> s0; stride = blk // EP
    g, kk = divmod(k, EP)
    j, i = divmod(vthread(cfg, p, t, g), stride)
    return j * blk + i + kk * stride


def warp_local(cfg, x):
    """exchange x (between passes x and x+1) stays inside each warp's block of 32*E coefficients"""
    N, NT, E = cfg[:3]
    return all(idx(cfg, p, t, k) // (32 * E) == t // 32 for p in (x, x + 1) for t in range(NT) for k in range(E))


def phys(cfg, x, i):
    A, S = cfg[5][x]
    return i + A * (i >> S) if A else i


def check(name):
    cfg = CONFIGS[name]
    N, NT, E, ns, eb, pads = cfg[:6]
    ok = True
    for x in range(len(pads)):
        # the header uses __syncwarp() for exchanges x >= 1 of a warp-blocked geometry: they must really be private to a warp
        assert warp_local(cfg, x) == (cfg[6] and x >= 1), (name, x)
        print(f"{name} exchange {x}: {'warp-private (__syncwarp)' if warp_local(cfg, x) else 'group barrier'}")
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


# ---- exchange hazards under arbitrary warp interleavings ------------------------------------------------------------------------
# With warp-private exchanges (__syncwarp) the warps of a group drift apart between group barriers.  This model replays the
# store / sync / load sequence of ntt.cuh (ntt_forward2s, ntt_inverse2s, ntt_forward, ntt_inverse with the alternating ExBuf) for
# every warp under random schedules that honour only the barriers the header places, with the physical addresses of the padded
# layouts, and checks that every load sees exactly the value its matching store wrote.
def _ops_fwd2s(cfg, tag):
    ops = [("B",)] if cfg[6] else []                              # pre_first_exchange
    for x in range(len(cfg[3]) - 1):
        s = "W" if warp_local(cfg, x) else "B"
        ops += [("st", "bx", x, x, (tag, "x", x)), (s,), ("ld", "bx", x, x + 1, (tag, "x", x)),
                ("st", "by", x, x, (tag, "y", x)), (s,), ("ld", "by", x, x + 1, (tag, "y", x))]
    return ops


def _ops_inv2s(cfg, tag):
    ops = []
    for x in reversed(range(len(cfg[3]) - 1)):
        s = "W" if warp_local(cfg, x) else "B"
        ops += [("st", "bx", x, x + 1, (tag, "x", x)), (s,), ("ld", "bx", x, x, (tag, "x", x)),
                ("st", "by", x, x + 1, (tag, "y", x)), (s,), ("ld", "by", x, x, (tag, "y", x))]
    return ops


def _ops_single(cfg, tag, inverse, state):
    """ntt_forward / ntt_inverse: successive exchanges alternate between the two buffers (ExBuf::next)"""
    ops = [("B",)] if (cfg[6] and not inverse) else []
    xs = range(len(cfg[3]) - 1)
    for x in (reversed(xs) if inverse else xs):
        buf = ("bx", "by")[state[0] & 1]; state[0] += 1
        s = "W" if warp_local(cfg, x) else "B"
        wp, rp = (x + 1, x) if inverse else (x, x + 1)
        ops += [("st", buf, x, wp, (tag, x)), (s,), ("ld", buf, x, rp, (tag, x))]
    return ops


def race_check(name, program, schedules=60, seed0=0):
    """program: list of 'F2' (forward2s), 'I2' (inverse2s), 'F' / 'I' (single transforms), 'B' (group barrier)"""
    import random
    cfg = CONFIGS[name]
    N, NT, E = cfg[:3]
    NW = NT // 32
    state = [0]
    ops = []
    for n, item in enumerate(program):
        ops += {"F2": lambda: _ops_fwd2s(cfg, n), "I2": lambda: _ops_inv2s(cfg, n), "F": lambda: _ops_single(cfg, n, False, state),
                "I": lambda: _ops_single(cfg, n, True, state), "B": lambda: [("B",)]}[item]()
    touched = {}

    def addrs(x, p, w):
        key = (x, p, w)
        if key not in touched:
            touched[key] = [phys(cfg, x, idx(cfg, p, t, k)) for t in range(32 * w, 32 * w + 32) for k in range(E)]
        return touched[key]
    for sched in range(schedules):
        rnd = random.Random(seed0 + sched)
        pc, waiting, mem = [0] * NW, [False] * NW, {"bx": {}, "by": {}}
        while any(p < len(ops) for p in pc):
            ready = [w for w in range(NW) if pc[w] < len(ops) and not waiting[w]]
            if not ready:
                raise AssertionError("deadlock")
            w = rnd.choice(ready); op = ops[pc[w]]
            if op[0] == "B":
                waiting[w] = True
                if all(waiting):
                    for v in range(NW):
                        waiting[v] = False; pc[v] += 1
                continue
            if op[0] == "W":
                pc[w] += 1; continue
            kind, buf, x, p, tag = op
            for a in addrs(x, p, w):
                if kind == "st":
                    mem[buf][a] = tag
                elif mem[buf].get(a) != tag:
                    return f"{name}: warp {w} loads {tag} from {buf}[{a}] and finds {mem[buf].get(a)} (schedule {sched})"
            pc[w] += 1
    return None
