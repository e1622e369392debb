#!/bin/bash
# Round-2 GPU call A (1 GPU): full GPU test suite, static-config kernel table, a short bench line, ncu --set full of the
# static-obstacle kernels (none existed).
O=gpurun_out/r02
mkdir -p $O
nvidia-smi > $O/gpu.txt 2>&1
python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
python tools/kernel_bench.py --configs 4obs_static,1obs_static,joint_limits,no_constraints > $O/kt_static_base.jsonl 2> $O/kt_static_base.err
python bench.py --steps 20 --warmup 3 > $O/bench_a.json 2> $O/bench_a.err; echo "bench exit $?" >> $O/bench_a.err
for spec in "planes cfg_joint_limits_and_4obs_static obs4_stat_planes" "rows cfg_joint_limits_and_4obs_static obs4_stat_rows" "planes cfg_joint_limits_and_1obs_static obs1_stat_planes"; do
  set -- $spec
  python tools/profile_target.py 24 $1 $2 > $O/plain_$3.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:lagrangian -s 2 -c 1 -o $O/ncu_$3 python tools/profile_target.py 24 $1 $2 > $O/ncu_$3.log 2>&1
done
tail -5 $O/pytest_gpu.log
