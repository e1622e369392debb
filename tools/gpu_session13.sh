#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|exit" gpurun_out/pytest_gpu.log | tail -10
timeout 600 python tools/kernel_bench.py --quick --libs libikl.so,libikl_cta.so,libikl.so,libikl_cta.so > gpurun_out/kt13.jsonl 2> gpurun_out/kt13.err; tail -2 gpurun_out/kt13.err
python - <<PY
import json
for l in open('gpurun_out/kt13.jsonl'):
    d=json.loads(l)
    if 'planes' in d['kernel']: print(d['lib'], d['kernel'][14:60], d['mode'], d['ms_best'], d['frac_of_hbm_peak'])
PY
