#!/bin/bash
# round 2, call AI: ncu (full set, source) of the proposal kernel of the final build
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"propose_kernel" -s 3 -c 1 -o gpurun_out/prof_ai python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/ai_ncu.log 2>&1
echo "ncu rc=$?"
