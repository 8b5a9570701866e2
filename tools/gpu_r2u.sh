#!/bin/bash
# round 2, call U: final single-GPU state: smoke, full GPU suite, bench line, launch list of one step
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/u_smoke.txt 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/u_smoke.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/u_pytest.txt 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/u_pytest.txt
timeout 600 python bench.py > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/u_bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['kernel_ms'], d['clocks'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/u_bench_ref.json 2> gpurun_out/u_bench_ref.err
echo "reference arm rc=$?"; cut -c1-300 gpurun_out/u_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/u_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-spectra > gpurun_out/u_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/u_launches.csv
