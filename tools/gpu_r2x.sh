#!/bin/bash
# round 2, call X: ncu (full set, source) of the two decay passes on the C3 + decays step
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"decay_kernel" -s 6 -c 2 -o gpurun_out/prof_x python bench.py --workload c3-decays --steps 1 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/x_ncu.log 2>&1
echo "ncu rc=$?"
