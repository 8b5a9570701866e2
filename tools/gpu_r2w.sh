#!/bin/bash
# round 2, call W: decay kernel with the species-only count pass: parity tests, C3 + decays bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sampler_gpu.py tests/test_facade_gpu.py tests/test_stats_gpu.py tests/test_chunk_gpu.py -q -x -k "decay or chunk or facade" > gpurun_out/w_pytest.txt 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/w_pytest.txt
timeout 600 python bench.py --workload c3-decays --no-cpu-baseline --no-spectra > gpurun_out/w_bench_c3-decays.json 2> gpurun_out/w_bench_c3-decays.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/w_bench_c3-decays.json')); print(d['ms_per_step'], d['value'], d['kernel_ms'], d.get('roofline_decay'))"
