#!/bin/bash
# round 2, call M: full GPU suite after the legacy sample files + guide build without memset, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/m_pytest.txt 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/m_pytest.txt
timeout 600 python bench.py > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err
echo "bench rc=$?"; cat gpurun_out/m_bench.json
