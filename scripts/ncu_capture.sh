#!/bin/bash
# ncu evidence for the hot-path kernels at a given batch (run under gpurun; one GPU).
#   scripts/ncu_capture.sh <tag> <batch> [full|dram] [kernel-regex] [extra driver args]
#   1. plain run of the driver (must exit 0), 2. launch list, 3. full: --set full + pipe / L2 / stall counters on K1..K6;
#      dram: only dram__bytes_{read,write}.sum (long launches, where the full set overflows its counters).
# Outputs: gpurun_out/<tag>_b<batch>.{plain.log,launches.csv,ncu-rep | dram.csv}; summarise with scripts/ncu_summary2.py, ncu_traffic.py.
set -u
TAG=$1; B=$2; MODE=${3:-full}; KREGEX=${4:-'l1_blind_rotate|l2_blind_rotate|trace_kernel|pack_kernel|keyswitch'}; shift; shift; shift || true; shift || true
OUT=gpurun_out/${TAG}_b${B}
mkdir -p gpurun_out
EXTRA="sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_alu.sum,\
sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active,\
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,lts__t_sectors_srcunit_tex.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,\
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum,\
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__average_warp_latency_issue_stalled_barrier.ratio,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,\
smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio,\
smsp__average_warp_latency_issue_stalled_mio_throttle.ratio,smsp__average_warp_latency_issue_stalled_wait.ratio,smsp__average_warp_latency_issue_stalled_not_selected.ratio,\
smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio,smsp__average_warp_latency_issue_stalled_lg_throttle.ratio,dram__bytes_read.sum,dram__bytes_write.sum"
CMD="python scripts/profile_driver.py --batch $B $*"
$CMD > ${OUT}.plain.log 2>&1 || { echo "plain run failed"; tail -5 ${OUT}.plain.log; exit 1; }
tail -1 ${OUT}.plain.log
if [ "$MODE" = "dram" ]; then
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"$KREGEX" -c 8 --csv --log-file ${OUT}.dram.csv $CMD > ${OUT}.ncu_dram.log 2>&1 || echo "ncu dram failed"
else
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${OUT}.launches.csv $CMD > ${OUT}.ncu_list.log 2>&1 || echo "launch list failed"
  ncu --set full --metrics "$EXTRA" --clock-control none --import-source on -k regex:"$KREGEX" -c 8 -f -o ${OUT} $CMD > ${OUT}.ncu_full.log 2>&1 || { echo "ncu full failed"; tail -5 ${OUT}.ncu_full.log; }
fi
ls -la ${OUT}*
