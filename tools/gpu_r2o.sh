#!/bin/bash
# round 2, call O: launch list (ncu, durations only) of the emulated 8-rank surface-chunk step
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/o_launches.csv \
    python tools/chunk_emulate.py --cells 1000000 --events 1000 --world 8 --steps 1 > gpurun_out/o_emulate.json 2> gpurun_out/o_emulate.err
echo "ncu rc=$?"; wc -l gpurun_out/o_launches.csv
