#!/bin/bash
# round 2, call I: full parity suite (legacy kind 0 included), final single-GPU numbers of all workloads
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_i.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_i.log
tail -6 gpurun_out/pytest_i.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_i.json'))
print('bench', d['value'], d['ms_per_step'], d['kernel_ms'], d['e2e']['value'], d['clocks'], d['cpu_baseline']['value'], d['cpu_baseline'].get('full_size_check',{}).get('cpu_seconds'))
PY
for w in c3-decays c3 c5; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --no-spectra --no-cpu-baseline > gpurun_out/bench_i_$w.json 2> gpurun_out/bench_i_$w.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_i_$w.json'))
    print('$w', d['value'], d['ms_per_step'], d['kernel_ms'], d['e2e']['value'])
except Exception as e:
    print('$w failed', e)
PY
done
timeout 600 python bench.py --impl reference --steps 1 > gpurun_out/bench_i_ref.json 2> gpurun_out/bench_i_ref.err
cut -c1-400 gpurun_out/bench_i_ref.json
python __graft_entry__.py > gpurun_out/smoke_i.log 2>&1; python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/smoke_i.log 2>&1; tail -2 gpurun_out/smoke_i.log
