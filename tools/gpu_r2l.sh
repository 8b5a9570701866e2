#!/bin/bash
# round 2, call L: ncu (full set, source) of the yield, prefix and QA kernels
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"yields_kernel|qa_kernel|tile_scan_kernel" -s 16 -c 4 -o gpurun_out/prof_l python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/l_ncu.log 2>&1
echo "ncu rc=$?"
