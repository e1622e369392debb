#!/bin/bash
# First GPU session: parity tests, smoke, bench, variant table, ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
timeout 900 python tools/kernel_bench.py --libs libikl.so > gpurun_out/kernel_table.jsonl 2> gpurun_out/kernel_table.err
timeout 600 python tools/kernel_bench.py --quick --libs libikl_v2.so,libikl_mb1.so,libikl_u1.so >> gpurun_out/kernel_table.jsonl 2>> gpurun_out/kernel_table.err
cat gpurun_out/kernel_table.jsonl | cut -c1-400
timeout 300 python tools/profile_target.py 24 > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_target.py 24 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lagrangian -s 2 -c 1 -o gpurun_out/prof_planes python tools/profile_target.py 24 planes > gpurun_out/ncu_full_planes.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lagrangian -s 2 -c 1 -o gpurun_out/prof_rows python tools/profile_target.py 24 rows > gpurun_out/ncu_full_rows.log 2>&1
tail -3 gpurun_out/ncu_full_planes.log
ls -la gpurun_out
