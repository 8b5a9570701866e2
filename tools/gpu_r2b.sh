#!/bin/bash
# round 2, call B: parity + timing of the reworked proposal kernel, ncu, host expansion probe
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_b.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_b.log
tail -8 gpurun_out/pytest_b.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-spectra --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_b.json'))
print(d['value'], d['ms_per_step'], d['kernel_ms'], d['e2e'], d['roofline']['fp64'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/b_ncu.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"propose_kernel|setup_kernel|bucket_kernel" -s 9 -c 3 -o gpurun_out/prof_b python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/b_ncu2.log 2>&1
echo "ncu full rc=$?"
g++ -O3 -march=native -o /tmp/expand_probe tools/expand_probe.cpp -lpthread && (nproc; /tmp/expand_probe 16; /tmp/expand_probe 8; /tmp/expand_probe 4) > gpurun_out/expand_probe.txt 2>&1
cat gpurun_out/expand_probe.txt
lscpu | head -20 > gpurun_out/lscpu.txt
