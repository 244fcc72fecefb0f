"""Per-kernel share of the step from an ncu launch list (scripts/ncu_capture.sh: <tag>.launches.csv).  Per-launch times under
ncu are cold-cache and serialised: compare SHARES with bench.py's CUDA-event stage times, not absolutes."""
import collections, csv, sys

for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]; kn, mv, mn, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        v = float(r[mv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu], 1e-6)
        a = agg.setdefault(r[kn].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{path}: {sum(a[0] for a in agg.values())} launches, {tot:.2f} ms")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        print(f"   {t:10.3f} ms {100 * t / tot:6.2f}%  x{c:<3d} {n[:100]}")
