#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|exit" gpurun_out/pytest_gpu.log | tail -30
timeout 600 python tools/kernel_bench.py --quick --libs libikl_t3.so,libikl_t4.so > gpurun_out/kernel_table2.jsonl 2> gpurun_out/kernel_table2.err
cut -c1-330 gpurun_out/kernel_table2.jsonl
