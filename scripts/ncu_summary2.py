"""Summarise every launch of an .ncu-rep captured by scripts/ncu_capture.sh: pipe utilisation (fp64 / fmaheavy / fmalite / alu),
issue slots, L2 and DRAM traffic, shared-memory bank conflicts, stall reasons, and the SASS opcode mix."""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
WANT = [
    ("time", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("block", "launch__block_size"), ("regs", "launch__registers_per_thread"),
    ("smem/CTA dyn", "launch__shared_mem_per_block_dynamic"), ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("eligible warps/clk", "smsp__warps_eligible.avg.per_cycle_active"),
    ("fp64 pipe %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), ("fma pipe %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("fmaheavy pipe %", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active"), ("fmalite pipe %", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active"),
    ("alu pipe %", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    ("inst fmaheavy", "sm__inst_executed_pipe_fmaheavy.sum"), ("inst fmalite", "sm__inst_executed_pipe_fmalite.sum"), ("inst fp64", "sm__inst_executed_pipe_fp64.sum"),
    ("inst alu", "sm__inst_executed_pipe_alu.sum"), ("inst total", "smsp__inst_executed.sum"),
    ("lsu wavefronts smem", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), ("lsu smem pipe %", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"), ("  ... loads", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
    ("  ... stores", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"),
    ("L2 bytes", "lts__t_bytes.sum"), ("L2 throughput %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("L1 hit %", "l1tex__t_sector_hit_rate.pct"), ("dram read", "dram__bytes_read.sum"), ("dram write", "dram__bytes_write.sum"),
    ("dram throughput %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("cycles", "sm__cycles_elapsed.avg"),
]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print("=" * 110)
    print(r[col["Kernel Name"]][:160])
    for label, m in WANT:
        if m in col:
            print(f"  {label:22s} {r[col[m]]:>24s} {units[col[m]]}")
    print("  stall cycles per issued instruction (smsp__average_warp_latency_issue_stalled_*):")
    for h, i in col.items():
        if h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio"):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v >= 0.05:
                print(f"      {h.replace('smsp__average_warp_latency_issue_stalled_', '').replace('.ratio', ''):28s} {v:8.2f}")
# opcode mix + stall samples per launch
nl = len(rows) - 2
for k in range(nl):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(k), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) < 4:
        continue
    h = srows[1]
    try:
        isrc, iex, ist = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    except ValueError:
        continue
    byop, stall, tot = collections.Counter(), collections.Counter(), 0
    for row in srows[2:]:
        try:
            n, s = int(row[iex]), int(row[ist])
        except (ValueError, IndexError):
            continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", row[isrc].strip())
        op = m.group(2) if m else row[isrc]
        op = ".".join(op.split(".")[:2]) if op.startswith(("IMAD", "LDS", "STS", "LDG", "LDL", "STL", "BAR", "ISETP")) else op.split(".")[0]
        byop[op] += n; stall[op] += s; tot += n
    print("-" * 110)
    print(f"launch {k}: {rows[2 + k][col['Kernel Name']][:90]}  executed warp-instructions {tot}")
    ss = max(1, sum(stall.values()))
    for op, n in byop.most_common(18):
        print(f"      {op:16s} {100 * n / max(1, tot):6.2f}% of instr   {100 * stall[op] / ss:6.2f}% of stall samples")
