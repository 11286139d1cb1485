#!/bin/bash
# Runs on the GPU box (gpurun): the ncu captures the round's profiles/ are built from, exported to small CSVs on the box
# (the .ncu-rep files themselves exceed what gpurun copies back).
#   scripts/gpu_round_profiles.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
# (1) every launch of one bench step with its time / DRAM bytes / tensor-pipe activity (kernel by kernel: P2P_GRAPH=0)
P2P_GRAPH=0 timeout 900 ncu --metrics $M --clock-control none --csv --log-file $OUT/${TAG}_step_metrics_ncu.csv python bench.py --profile > $OUT/${TAG}_ncu_step.log 2>&1
# (2) one 256-crop forward, full set: DRAM traffic + per-layer table
timeout 1000 ncu --set full --clock-control none -s 70 -c 35 -f -o /tmp/${TAG}_fwd256 python scripts/profile_forward.py 256 > $OUT/${TAG}_ncu_fwd.log 2>&1
ncu -i /tmp/${TAG}_fwd256.ncu-rep --page raw --csv > $OUT/${TAG}_fwd256_ncu_raw.csv 2>/dev/null
# (3) the RANSAC hypothesis kernel, full set + source-level stall samples
P2P_GRAPH=0 timeout 500 ncu --set full --clock-control none --import-source on -k regex:ransac_hyp -s 1 -c 1 -f -o /tmp/${TAG}_hyp python bench.py --profile > $OUT/${TAG}_ncu_hyp.log 2>&1
ncu -i /tmp/${TAG}_hyp.ncu-rep --page raw --csv > $OUT/${TAG}_hyp_ncu_raw.csv 2>/dev/null
ls -la $OUT | grep ${TAG}_
