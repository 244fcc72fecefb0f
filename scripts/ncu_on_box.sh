#!/bin/bash
# Capture + summarise on the GPU box, keep only the text (the .ncu-rep files of a full capture exceed gpurun's 64 MiB return limit).
#   scripts/ncu_on_box.sh <tag>      -> gpurun_out/<tag>_ncu_batch{296,2368}.txt, <tag>_launch_shares.txt, profiles-style traffic JSON
set -u
TAG=$1
bash scripts/ncu_capture.sh $TAG 296 full
python scripts/ncu_summary2.py gpurun_out/${TAG}_b296.ncu-rep > gpurun_out/${TAG}_ncu_batch296.txt 2>&1
bash scripts/ncu_capture.sh $TAG 2368 full "l1_blind_rotate|l2_blind_rotate|keyswitch_dp4a|trace_kernel"
python scripts/ncu_summary2.py gpurun_out/${TAG}_b2368.ncu-rep > gpurun_out/${TAG}_ncu_batch2368.txt 2>&1
bash scripts/ncu_capture.sh $TAG 16384 dram "l1_blind_rotate|l2_blind_rotate|trace_kernel|keyswitch_dp4a|pack_kernel"
python scripts/launch_shares.py gpurun_out/${TAG}_b296.launches.csv gpurun_out/${TAG}_b2368.launches.csv > gpurun_out/${TAG}_launch_shares.txt
python scripts/ncu_traffic.py ${TAG}box gpurun_out/${TAG}_b296.ncu-rep:296 gpurun_out/${TAG}_b2368.ncu-rep:2368 gpurun_out/${TAG}_b16384.dram.csv:16384
cp profiles/${TAG}box_dram_traffic.json gpurun_out/${TAG}_dram_traffic.json
rm -f gpurun_out/${TAG}_b296.ncu-rep gpurun_out/${TAG}_b2368.ncu-rep
ls -la gpurun_out/
