#!/bin/bash
# round 2, call C: parity, CTA-size variants, other workloads, ncu of decay / legacy kernels, e2e phases,
# full-size reference run
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_c.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_c.log
tail -5 gpurun_out/pytest_c.log
for t in 768 640 512; do
  ISS_SAMPLER_THREADS=$t timeout 600 python bench.py --steps 10 --warmup 3 --no-spectra --no-cpu-baseline > gpurun_out/bench_c_t$t.json 2> gpurun_out/bench_c_t$t.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_c_t$t.json'))
print('threads $t', d['value'], d['ms_per_step'], d['kernel_ms'], d['e2e']['value'])
PY
done
for w in c3-decays c3 c5; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --no-spectra --no-cpu-baseline > gpurun_out/bench_c_$w.json 2> gpurun_out/bench_c_$w.err
  echo "bench $w rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_c_$w.json'))
    print('$w', d['value'], d['ms_per_step'], d['kernel_ms'], d['e2e']['value'], d.get('roofline_decay'))
except Exception as e:
    print('$w failed', e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"decay_kernel" -s 6 -c 2 -o gpurun_out/prof_c_decay python bench.py --workload c3-decays --steps 1 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/c_ncu_decay.log 2>&1
echo "ncu decay rc=$?"
NEV=100 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"legacy_sample_kernel" -s 1 -c 1 -o gpurun_out/prof_c_legacy python tools/legacy_probe.py > gpurun_out/c_ncu_legacy.log 2>&1
echo "ncu legacy rc=$?"
E2E_CALLS=4 timeout 600 python tools/e2e_probe.py > gpurun_out/e2e_probe_c.txt 2>&1
grep "profile\|===" gpurun_out/e2e_probe_c.txt | tail -40
timeout 900 python bench.py --impl reference --full-size --steps 1 > gpurun_out/ref_fullsize_c4.json 2> gpurun_out/ref_fullsize_c4.err
echo "ref fullsize rc=$?"
cat gpurun_out/ref_fullsize_c4.json | cut -c1-1500
