#!/bin/bash
# round 2, call P (2 GPUs): multi-GPU tests and the surface-chunk probe with the collectives behind the C ABI
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -x > gpurun_out/p_pytest.txt 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/p_pytest.txt
for c in abi torch; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tools/chunk_probe.py --cells 1000000 --events 1000 --collectives $c > gpurun_out/p_chunk_n2_$c.json 2> gpurun_out/p_chunk_n2_$c.err
echo "chunk_probe $c rc=$?"; cat gpurun_out/p_chunk_n2_$c.json
done
