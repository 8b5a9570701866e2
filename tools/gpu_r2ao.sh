#!/bin/bash
# round 2, call AO: launch list (durations) of the C4 step, final build
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 110 --csv --log-file gpurun_out/ao_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-spectra > gpurun_out/ao_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/ao_launches.csv
