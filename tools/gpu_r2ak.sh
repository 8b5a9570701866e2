#!/bin/bash
# round 2, call AK: short bench line (kernel times) used for one-line kernel variants
mkdir -p gpurun_out; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-spectra > gpurun_out/y_b.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/y_b.json')); print(round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['kernel_ms'].items()})"
