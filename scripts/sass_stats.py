"""Static SASS statistics of one kernel of libomr_b200.so: instruction count, register moves, and the loops (backward
branches) with their body sizes.  Usage: python scripts/sass_stats.py <substring of the mangled kernel name> [lib]"""
import re, subprocess, sys, collections
lib = sys.argv[2] if len(sys.argv) > 2 else "tfhe-omr_b200/lib/libomr_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s+Function : ", txt)
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if sys.argv[1] not in name:
        continue
    ins = [(int(m.group(1), 16), m.group(2)) for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", b)]
    ops = collections.Counter()
    for a, t in ins:
        p = t.split()
        ops[(p[1] if p[0].startswith("@") else p[0])] += 1
    mv = sum(v for k, v in ops.items() if k.startswith("MOV") or k.startswith("IMAD.MOV"))
    print(name, "instrs", len(ins), "moves", mv)
    for a, t in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s+)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            body = [x for x in ins if lo <= x[0] <= a]
            bm = sum(1 for x in body if re.match(r"(@\S+\s+)?(MOV|IMAD\.MOV)", x[1]))
            print(f"  loop {lo:#x}..{a:#x}: {len(body)} instrs, {bm} moves")
    print("  top:", ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
