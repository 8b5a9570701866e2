#!/bin/bash
# round 2, call F (8 GPUs): multi-GPU tests, concurrent D2H ceiling, weak scaling of the bench, chunk probe
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_f.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q > gpurun_out/pytest_f.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_f.log
tail -6 gpurun_out/pytest_f.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29540+n)) tools/d2h_probe.py --gb 2.2 2>/dev/null | grep concurrent_d2h >> gpurun_out/d2h_probe_f.jsonl
done
cat gpurun_out/d2h_probe_f.jsonl
for n in 8 4 2; do
  timeout 600 $TR --nproc-per-node $n --master-port $((29560+n)) bench.py --gpus $n --steps 10 --warmup 3 --no-spectra > gpurun_out/bench_f_n$n.json 2> gpurun_out/bench_f_n$n.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_f_n$n.json'))
    print('N=$n', d['value'], d['ms_per_step'], d['e2e']['value'])
except Exception as e:
    print('N=$n failed', e)
PY
done
timeout 600 $TR --nproc-per-node 8 --master-port 29590 tools/chunk_probe.py --cells 1000000 --events 1000 --steps 3 > gpurun_out/chunk_probe_f_8gpu.json 2> gpurun_out/chunk_probe_f.err
cat gpurun_out/chunk_probe_f_8gpu.json | cut -c1-1200
