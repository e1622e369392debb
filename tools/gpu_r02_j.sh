#!/bin/bash
# Round-2 GPU call J (1 GPU): evidence run with the final kernels of the round (gradient-only training launches, peer-memory
# gradient all-reduce): full GPU suite, smoke, complete kernel table, bench line + reference arm, ncu launch list of the bench
# command, ncu --set full of the benchmark kernel and of the gradient-only static-obstacle kernels.
O=gpurun_out/r02j
mkdir -p $O
python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu_final.log 2>&1; echo "pytest exit $? ($(date -u +%Y-%m-%dT%H:%M:%SZ))" >> $O/pytest_gpu_final.log
tail -4 $O/pytest_gpu_final.log
python __graft_entry__.py smoke > $O/smoke_final.log 2>&1; echo "smoke exit $?" >> $O/smoke_final.log; tail -2 $O/smoke_final.log
python tools/parity_report.py --out $O/parity_report_final.json > $O/parity_report_final.log 2>&1; echo "parity exit $?" >> $O/parity_report_final.log
python tools/kernel_bench.py --iters 30 > $O/kernel_table_final.jsonl 2> $O/kernel_table_final.err
python bench.py --steps 20 --warmup 3 > $O/bench_final.json 2> $O/bench_final.err; echo "bench exit $?" >> $O/bench_final.err
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python bench.py --steps 20 --warmup 3 --skip-e2e --skip-train-step --skip-cpu --skip-context --skip-strong > $O/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_final.csv \
  python bench.py --steps 20 --warmup 3 --skip-e2e --skip-train-step --skip-cpu --skip-context --skip-strong > $O/ncu_bench.log 2>&1
for spec in "planes cfg_joint_limits_and_4obs_dynamic fwdbwd dyn4_planes" "planes cfg_joint_limits_and_4obs_static grad obs4_stat_planes_gradonly" "rows cfg_joint_limits_and_4obs_static grad obs4_stat_rows_gradonly" "planes cfg_joint_limits_and_1obs_static grad obs1_stat_planes_gradonly"; do
  set -- $spec
  python tools/profile_target.py 24 $1 $2 $3 > $O/plain_$4.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:lagrangian -s 2 -c 1 -o $O/ncu_$4 python tools/profile_target.py 24 $1 $2 $3 > $O/ncu_$4.log 2>&1
  # gpurun brings back at most 64 MiB: summarise on the box (details page, trimmed raw metrics, SASS-level source page), drop the report
  python tools/ncu_summarise.py $O/ncu_$4.ncu-rep $O/ncu_$4 > /dev/null 2>&1
  ncu -i $O/ncu_$4.ncu-rep --page source --csv 2> /dev/null | gzip -9 > $O/ncu_$4.source.csv.gz
  rm -f $O/ncu_$4.ncu-rep
done
ls -la $O
