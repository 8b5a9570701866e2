#!/bin/bash
# round 2, call AC: final state: smoke, full GPU suite, default bench line, C3 + decays line
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/ac_smoke.txt 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/ac_smoke.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/ac_pytest.txt 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/ac_pytest.txt
timeout 600 python bench.py > gpurun_out/ac_bench.json 2> gpurun_out/ac_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/ac_bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['d2h_link'], d['kernel_ms'])"
timeout 600 python bench.py --workload c3-decays --no-cpu-baseline --no-spectra > gpurun_out/ac_bench_c3-decays.json 2> gpurun_out/ac_bench_c3-decays.err
echo "bench c3-decays rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/ac_bench_c3-decays.json')); print(d['ms_per_step'], d['value'], d['kernel_ms'])"
