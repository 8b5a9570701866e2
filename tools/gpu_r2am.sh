#!/bin/bash
# round 2, call AM: end-to-end call time against the number of event batches (ISS_BATCHES) on this box
mkdir -p gpurun_out
for b in 4 8 16 32; do
ISS_BATCHES=$b timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/am_bench_$b.json 2> gpurun_out/am_bench_$b.err
python -c "
import json; d=json.load(open('gpurun_out/am_bench_$b.json')); e=d['e2e']; print('batches $b', round(e['d2h_link']['measured_ms_per_call'],2), 'ms per call, link', round(e['d2h_link']['gbs_per_rank_all_ranks_busy'],1), 'GB/s, floor', round(e['d2h_link']['floor_ms_per_call'],1))"
done
