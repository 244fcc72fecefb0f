"""Summarise an .ncu-rep: headline raw metrics + executed-instruction histogram by opcode (needs -lineinfo/--import-source)."""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
skip = sys.argv[2] if len(sys.argv) > 2 else "0"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg']
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f"{w:75s} {r[i]:>22s} {units[i]}")
for i, h in enumerate(hdr):
    if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
        try:
            v = float(r[i])
        except ValueError:
            continue
        if v > 2.0:
            print(f"  stall {h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''):40s} {v:8.2f} %")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 3:
    hdr = rows[1]
    isrc, iex, ist = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
    byop, stall, tot = collections.Counter(), collections.Counter(), 0
    for row in rows[2:]:
        try:
            n, s = int(row[iex]), int(row[ist])
        except (ValueError, IndexError):
            continue
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', row[isrc].strip())
        op = m.group(2) if m else row[isrc]
        op = '.'.join(op.split('.')[:2]) if op.startswith(('IMAD', 'LDS', 'STS', 'LDG', 'LDL', 'STL', 'BAR', 'ISETP')) else op.split('.')[0]
        byop[op] += n; stall[op] += s; tot += n
    print(f"executed warp-instructions {tot}")
    ss = max(1, sum(stall.values()))
    for op, n in byop.most_common(26):
        print(f"  {op:16s} {100 * n / tot:6.2f}% of instr   {100 * stall[op] / ss:6.2f}% of stall samples")
