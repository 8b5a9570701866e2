#!/bin/bash
# round 2, call D: partition kernel (cell blocks, tile-level reservations), e2e call-time distribution
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_d.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_d.log
tail -5 gpurun_out/pytest_d.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-spectra --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_d.json'))
print('bench', d['value'], d['ms_per_step'], d['kernel_ms'], d['e2e']['value'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_d.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/d_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"propose_kernel|setup_kernel|partition_kernel" -s 9 -c 3 -o gpurun_out/prof_d python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/d_ncu2.log 2>&1
echo "ncu full rc=$?"
E2E_CALLS=30 timeout 600 python tools/e2e_probe.py > gpurun_out/e2e_probe_d.txt 2>&1
grep "===" gpurun_out/e2e_probe_d.txt | tr '\n' ' '
