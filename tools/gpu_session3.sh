#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|exit" gpurun_out/pytest_gpu.log | tail -30
timeout 900 python tools/kernel_bench.py --libs libikl.so > gpurun_out/kernel_table3.jsonl 2> gpurun_out/kernel_table3.err
python - <<'PY'
import json
for l in open('gpurun_out/kernel_table3.jsonl'):
    d=json.loads(l); print(d['config'][4:], d['layout'], d['mode'], d['ms_best'], d['frac_of_hbm_peak'])
PY
timeout 300 python tools/profile_target.py 24 planes > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lagrangian -s 2 -c 1 -o gpurun_out/prof_planes_packed python tools/profile_target.py 24 planes > gpurun_out/ncu_full_planes.log 2>&1
tail -2 gpurun_out/ncu_full_planes.log
