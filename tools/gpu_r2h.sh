#!/bin/bash
# round 2, call H: staged partition kernel; ncu of the sampler kernels for the record
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_h.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_h.log
tail -6 gpurun_out/pytest_h.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-spectra --no-cpu-baseline > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_h.json'))
print('bench', d['value'], d['ms_per_step'], d['kernel_ms'], d['e2e']['value'], d['clocks'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file gpurun_out/launches_h.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/h_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_h.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki].split('(')[0][:60], float(r[vi].replace(',',''))) for r in rows[hi+2:] if len(r)>vi]
idx=[i for i,(n,v) in enumerate(seq) if 'propose' in n and v>5e6]
for n,v in seq[idx[-1]-8:idx[-1]+3]: print("%-62s %10.1f us"%(n,v/1000))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"propose_kernel|setup_kernel|partition_kernel|qa_kernel|guide_kernel" -s 15 -c 5 -o gpurun_out/prof_h python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/h_ncu2.log 2>&1
echo "ncu full rc=$?"
