#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|exit" gpurun_out/pytest_gpu.log | tail -10
for tma in 1 0; do
IKL_PLANES_TMA=$tma timeout 300 python tools/kernel_bench.py --quick > gpurun_out/kt_tma$tma.jsonl 2> gpurun_out/kt_tma$tma.err; tail -2 gpurun_out/kt_tma$tma.err
python - <<PY
import json
for l in open('gpurun_out/kt_tma$tma.jsonl'):
    d=json.loads(l); print('tma=$tma', d['kernel'][14:60], d['mode'], d['ms_best'], d['frac_of_hbm_peak'])
PY
done
