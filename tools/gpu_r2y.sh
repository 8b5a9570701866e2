#!/bin/bash
# round 2, call Y: boost-invariant specialisation of the proposal kernel: parity (2+1D cases) and the C3 lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sampler_gpu.py tests/test_stats_gpu.py tests/test_facade_gpu.py -q -x > gpurun_out/y_pytest.txt 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/y_pytest.txt
for w in c3 c3-decays; do
timeout 600 python bench.py --workload $w --no-cpu-baseline --no-spectra > gpurun_out/y_bench_$w.json 2> gpurun_out/y_bench_$w.err
echo "bench $w rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/y_bench_$w.json')); print(d['ms_per_step'], d['value'], d['kernel_ms'])"
done
