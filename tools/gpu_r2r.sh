#!/bin/bash
# round 2, call R: ncu (full set, source) of the QA kernel on the C4 step; e2e phases + D2H probe of this box
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"qa_kernel" -s 2 -c 1 -o gpurun_out/prof_r python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/r_ncu.log 2>&1
echo "ncu rc=$?"
ISS_PROFILE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r_bench.json')); print(d['ms_per_step'], d['e2e'], d['kernel_ms'])"
grep "iss profile\|generate_samples" gpurun_out/r_bench.err | tail -14
timeout 300 python tools/d2h_probe.py > gpurun_out/r_d2h.jsonl 2>&1; tail -3 gpurun_out/r_d2h.jsonl
