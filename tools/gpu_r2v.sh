#!/bin/bash
# round 2, call V (8 GPUs): final weak-scaling bench line and the multi-GPU tests
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/v_bench_n8.json 2> gpurun_out/v_bench_n8.err
echo "bench n8 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/v_bench_n8.json')); print(d['ms_per_step'], d['value'], d['e2e'], d['ms_per_step_per_rank'], d['clocks'])"
timeout 600 python -m pytest tests/test_multigpu_gpu.py -q > gpurun_out/v_pytest_multigpu.txt 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/v_pytest_multigpu.txt
