"""DRAM traffic per launch of the hot-path kernels from ncu captures -> profiles/<tag>_dram_traffic.json, the file bench.py reads
for `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum, one launch).
    python scripts/ncu_traffic.py <tag> <rep-or-csv>:<batch> [...]
A .ncu-rep (any section set that holds the dram counters) or the --csv log of a `--metrics dram__bytes_read.sum,dram__bytes_write.sum` run."""
import csv, io, json, os, subprocess, sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows_of(path):
    if path.endswith(".ncu-rep"):
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        h, u = rows[0], rows[1]
        for r in rows[2:]:
            if len(r) < len(h):
                continue
            out = {"kernel": r[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("omr::", "")}
            for key, m in (("dram_read_bytes", "dram__bytes_read.sum"), ("dram_write_bytes", "dram__bytes_write.sum")):
                try:
                    out[key] = float(r[h.index(m)]) * UNIT.get(u[h.index(m)], 1)
                except ValueError:
                    out[key] = None
            yield out
    else:                                     # ncu --csv --log-file of a plain metrics run: one row per (launch, metric)
        rows = list(csv.reader(open(path)))
        hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
        h = rows[hi]; acc = {}
        for r in rows[hi + 1:]:
            if len(r) < len(h):
                continue
            k = (r[h.index("ID")], r[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("omr::", ""))
            v = float(r[h.index("Metric Value")].replace(",", "")) * UNIT.get(r[h.index("Metric Unit")], 1)
            acc.setdefault(k, {})[r[h.index("Metric Name")]] = v
        for (_, name), m in acc.items():
            yield {"kernel": name, "dram_read_bytes": m.get("dram__bytes_read.sum"), "dram_write_bytes": m.get("dram__bytes_write.sum")}


if __name__ == "__main__":
    tag, out = sys.argv[1], []
    for spec in sys.argv[2:]:
        path, batch = spec.rsplit(":", 1)
        for row in rows_of(path):
            if row["dram_read_bytes"] is None or row["dram_read_bytes"] != row["dram_read_bytes"]:
                continue
            row.update(batch=int(batch), source=os.path.basename(path)); out.append(row)
    dst = os.path.join(ROOT, "profiles", f"{tag}_dram_traffic.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(f"wrote {dst}: {len(out)} rows")
