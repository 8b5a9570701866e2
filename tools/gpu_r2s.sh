#!/bin/bash
# round 2, call S: ncu (full set, source) of the set-up and QA kernels on the C4 step
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"setup_kernel|qa_kernel" -s 4 -c 2 -o gpurun_out/prof_s python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/s_ncu.log 2>&1
echo "ncu rc=$?"
