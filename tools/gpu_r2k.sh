#!/bin/bash
# round 2, call K: QA kernel with edge-based binning
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_k.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_k.log
tail -6 gpurun_out/pytest_k.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-spectra --no-cpu-baseline > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_k.json'))
print('bench', d['value'], d['ms_per_step'], d['kernel_ms'], d['e2e']['value'], d['clocks'])
PY
