#!/bin/bash
# round 2, call A: parity of the cell-sorted sampler + first timing + ncu of the proposal kernel
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_a.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_a.log
tail -15 gpurun_out/pytest_a.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-spectra > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
echo "bench rc=$?"
cat gpurun_out/bench_a.json | head -c 3000
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_a.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/b_ncu.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:propose_kernel -s 3 -c 1 -o gpurun_out/prof_propose_a python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/b_ncu2.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
