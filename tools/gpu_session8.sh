#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|exit" gpurun_out/pytest_gpu.log | tail -10
timeout 600 python tools/kernel_bench.py --quick > gpurun_out/kt8.jsonl 2> gpurun_out/kt8.err; tail -2 gpurun_out/kt8.err
python - <<PY
import json
for l in open('gpurun_out/kt8.jsonl'):
    d=json.loads(l); print(d['lib'], d['kernel'][14:60], d['mode'], d['ms_best'], d['frac_of_hbm_peak'])
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ragged or golden or tie or side_outputs or dual_update" > gpurun_out/memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/memcheck.log
grep -E "ERROR SUMMARY|passed|failed|memcheck exit|Invalid|========= " gpurun_out/memcheck.log | head -20
