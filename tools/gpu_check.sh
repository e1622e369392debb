#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, kernel table and (optionally) the ncu captures.
#   gpurun --timeout 2400 -- 'bash tools/gpu_check.sh [ncu]'
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|exit" gpurun_out/pytest_gpu.log | tail -12
grep -E "^E  " gpurun_out/pytest_gpu.log | head -20
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 900 python tools/kernel_bench.py > gpurun_out/kernel_table.jsonl 2> gpurun_out/kernel_table.err
if [ "$1" = "ncu" ]; then
  timeout 300 python tools/profile_target.py 24 planes > gpurun_out/plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:lagrangian -s 2 -c 1 -o gpurun_out/prof_planes python tools/profile_target.py 24 planes > gpurun_out/ncu_full.log 2>&1
  timeout 300 python tools/profile_target.py 24 rows > gpurun_out/plain_rows.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:lagrangian -s 2 -c 1 -o gpurun_out/prof_rows python tools/profile_target.py 24 rows > gpurun_out/ncu_full_rows.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 20 --warmup 3 --skip-e2e --skip-cpu --skip-train-step --skip-context > gpurun_out/ncu_bench.log 2>&1
fi
