#!/bin/bash
# round 2, call AF: DRAM bytes and duration of every launch of the C4 step (whole-step traffic)
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/af_traffic.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-spectra > gpurun_out/af_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/af_traffic.csv
