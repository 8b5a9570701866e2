#!/bin/bash
# round 2, call Z (N GPUs): weak-scaling bench line of the final build
mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/z_bench_n$N.json 2> gpurun_out/z_bench_n$N.err
echo "bench n$N rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/z_bench_n$N.json')); print(d['ms_per_step'], d['value'], d['e2e'], d['ms_per_step_per_rank'])"
