#!/bin/bash
# Round-2 GPU call E (1 GPU): evidence run with the final kernels -- full GPU suite, smoke, parity report at 2^24, complete
# kernel table, bench line, ncu launch list of the bench command and ncu --set full of the benchmark kernel and the
# static-obstacle kernel.
O=gpurun_out/r02
mkdir -p $O
python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu_final.log 2>&1; echo "pytest exit $? ($(date -u +%Y-%m-%dT%H:%M:%SZ))" >> $O/pytest_gpu_final.log
tail -4 $O/pytest_gpu_final.log
python __graft_entry__.py smoke > $O/smoke_final.log 2>&1; echo "smoke exit $?" >> $O/smoke_final.log; tail -2 $O/smoke_final.log
python tools/parity_report.py --out $O/parity_report_final.json > $O/parity_report_final.log 2>&1; echo "parity exit $?" >> $O/parity_report_final.log
python tools/kernel_bench.py --iters 30 > $O/kernel_table_final.jsonl 2> $O/kernel_table_final.err
python bench.py --steps 20 --warmup 3 > $O/bench_final.json 2> $O/bench_final.err; echo "bench exit $?" >> $O/bench_final.err
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
# ncu: launch list of the bench command (skip the side legs), then full captures
python bench.py --steps 20 --warmup 3 --skip-e2e --skip-train-step --skip-cpu --skip-context --skip-strong > $O/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_final.csv \
  python bench.py --steps 20 --warmup 3 --skip-e2e --skip-train-step --skip-cpu --skip-context --skip-strong > $O/ncu_bench.log 2>&1
for spec in "planes cfg_joint_limits_and_4obs_dynamic dyn4_planes" "rows cfg_joint_limits_and_4obs_dynamic dyn4_rows" "planes cfg_joint_limits_and_4obs_static obs4_stat_planes_final"; do
  set -- $spec
  python tools/profile_target.py 24 $1 $2 > $O/plain_$3.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:lagrangian -s 2 -c 1 -o $O/ncu_$3 python tools/profile_target.py 24 $1 $2 > $O/ncu_$3.log 2>&1
done
