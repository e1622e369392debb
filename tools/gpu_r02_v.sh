#!/bin/bash
# Round-2 GPU call V (1 GPU): evidence run with the final build (dual step closed inside its last launch, ABI 5): full GPU suite,
# smoke, one complete bench line + the reference arm, ncu launch list of the bench command.
O=gpurun_out/r02v
mkdir -p $O
python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu_final.log 2>&1; echo "pytest exit $? ($(date -u +%Y-%m-%dT%H:%M:%SZ))" >> $O/pytest_gpu_final.log
tail -4 $O/pytest_gpu_final.log
python __graft_entry__.py smoke > $O/smoke_final.log 2>&1; echo "smoke exit $?" >> $O/smoke_final.log; tail -2 $O/smoke_final.log
python bench.py --steps 20 --warmup 3 > $O/bench_final.json 2> $O/bench_final.err; echo "bench exit $?" >> $O/bench_final.err
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python bench.py --steps 20 --warmup 3 --skip-e2e --skip-train-step --skip-cpu --skip-context --skip-strong > $O/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_final.csv \
  python bench.py --steps 20 --warmup 3 --skip-e2e --skip-train-step --skip-cpu --skip-context --skip-strong > $O/ncu_bench.log 2>&1
python tools/kernel_bench.py --iters 30 > $O/kernel_table_final.jsonl 2> $O/kernel_table_final.err
ls -la $O; head -c 600 $O/bench_final.json
