#!/bin/bash
# round 2, call AJ: ten-term series of the |p| sampler as Horner polynomials: parity, bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sampler_gpu.py tests/test_chunk_gpu.py tests/test_stats_gpu.py tests/test_facade_gpu.py -q -x > gpurun_out/aj_pytest.txt 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/aj_pytest.txt
timeout 600 python bench.py --no-cpu-baseline --no-spectra > gpurun_out/aj_bench.json 2> gpurun_out/aj_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/aj_bench.json')); print(d['ms_per_step'], d['value'], d['kernel_ms'])"
