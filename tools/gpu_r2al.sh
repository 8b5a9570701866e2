#!/bin/bash
# round 2, call AL: per-call times and host phases of the end-to-end loop (10 calls)
mkdir -p gpurun_out
ISS_BENCH_TRACE=1 ISS_PROFILE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/al_bench.json 2> gpurun_out/al_bench.err
echo "rc=$?"; grep -c "generate_samples" gpurun_out/al_bench.err; grep "=== generate_samples\|shell total\|batches (sample\|compute_yields\|final fetch" gpurun_out/al_bench.err | tail -60
