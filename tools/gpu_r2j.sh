#!/bin/bash
# round 2, call J (8 GPUs): weak scaling with the QA all-reduce through the C ABI, chunk probe
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 2; do
  timeout 600 $TR --nproc-per-node $n --master-port $((29660+n)) bench.py --gpus $n --steps 10 --warmup 3 --no-spectra > gpurun_out/bench_j_n$n.json 2> gpurun_out/bench_j_n$n.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_j_n$n.json'))
    print('N=$n', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['qa_allreduce'], d['ms_per_step_per_rank'], d['clocks'])
except Exception as e:
    print('N=$n failed', e)
PY
done
ISS_BENCH_TRACE=1 timeout 600 $TR --nproc-per-node 8 --master-port 29680 bench.py --gpus 8 --steps 10 --warmup 3 --no-spectra > gpurun_out/bench_j_n8b.json 2> gpurun_out/bench_j_n8b.err
grep "bench trace" gpurun_out/bench_j_n8b.err | sort | awk '{print $4, $7}' | tr '\n' ' ' | cut -c1-1500
python - <<PY
import json
d=json.load(open('gpurun_out/bench_j_n8b.json'))
print('N=8 second run', d['value'], d['ms_per_step'], d['e2e']['value'])
PY
timeout 600 $TR --nproc-per-node 8 --master-port 29690 tools/chunk_probe.py --cells 1000000 --events 1000 --steps 3 > gpurun_out/chunk_probe_j_8gpu.json 2> gpurun_out/chunk_probe_j.err
cut -c1-900 gpurun_out/chunk_probe_j_8gpu.json
