#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|exit" gpurun_out/pytest_gpu.log | tail -12
grep -E "^E  " gpurun_out/pytest_gpu.log | head -20
timeout 900 python bench.py --steps 100 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline_c'])"
