#!/bin/bash
# round 2, call N: surface-chunk mode with the two-comparison ownership passes: identity tests, then
# the ranks of an 8-rank run emulated on one GPU (even and balanced cut)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chunk_gpu.py tests/test_multigpu_gpu.py -q -x > gpurun_out/n_pytest.txt 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/n_pytest.txt
timeout 900 python tools/chunk_emulate.py --cells 1000000 --events 1000 --world 8 > gpurun_out/n_emulate.json 2> gpurun_out/n_emulate.err
echo "emulate rc=$?"; tail -c 3000 gpurun_out/n_emulate.json; tail -3 gpurun_out/n_emulate.err
