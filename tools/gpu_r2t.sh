#!/bin/bash
# round 2, call T: variants of the cell search of setup_kernel (ISS_SETUP_TUNE), parity + timing.
# Kept for the record: all variants were slower (history) and were removed from
# sampler.cu after this call, so ISS_SETUP_TUNE has no effect on the committed code.
mkdir -p gpurun_out
for t in 4 5; do
ISS_SETUP_TUNE=$t timeout 600 python -m pytest tests/test_sampler_gpu.py tests/test_chunk_gpu.py -q -x > gpurun_out/t_pytest_$t.txt 2>&1
echo "tune $t pytest rc=$?"; tail -1 gpurun_out/t_pytest_$t.txt
done
for t in 0 1 2 4 5; do
ISS_SETUP_TUNE=$t timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/t_bench_$t.json 2> gpurun_out/t_bench_$t.err
python -c "
import json; d=json.load(open('gpurun_out/t_bench_$t.json')); print('tune $t', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['kernel_ms'].items()})"
done
