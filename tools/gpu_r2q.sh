#!/bin/bash
# round 2, call Q (8 GPUs): surface-chunk probe, C4 and C5, collectives behind the C ABI
mkdir -p gpurun_out
run() {  # name, timeout, args...
    local name=$1 to=$2; shift 2
    timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/chunk_probe.py "$@" > gpurun_out/q_chunk_$name.json 2> gpurun_out/q_chunk_$name.err
    echo "chunk_probe $name rc=$?"; cat gpurun_out/q_chunk_$name.json
}
run c4 300 --cells 1000000 --events 1000 --steps 5
run c5 420 --cells 10000000 --events 100 --steps 3
