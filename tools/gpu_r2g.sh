#!/bin/bash
# round 2, call G: guide table of the cell search, C5-size test, timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_g.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_g.log
tail -6 gpurun_out/pytest_g.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-spectra --no-cpu-baseline > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_g.json'))
print('bench', d['value'], d['ms_per_step'], d['kernel_ms'], d['e2e']['value'], d['clocks'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file gpurun_out/launches_g.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/g_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_g.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki].split('(')[0][:60], float(r[vi].replace(',',''))) for r in rows[hi+2:] if len(r)>vi]
idx=[i for i,(n,v) in enumerate(seq) if 'propose' in n and v>5e6]
for n,v in seq[idx[-1]-32:idx[-1]+3]: print("%-62s %10.1f us"%(n,v/1000))
PY
