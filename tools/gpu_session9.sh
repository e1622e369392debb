#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_target.py 24 planes > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lagrangian -s 2 -c 1 -o gpurun_out/prof_planes_tma python tools/profile_target.py 24 planes > gpurun_out/ncu_full_planes.log 2>&1
tail -2 gpurun_out/ncu_full_planes.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 20 --warmup 3 --skip-e2e --skip-cpu --skip-train-step > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
