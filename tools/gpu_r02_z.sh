#!/bin/bash
# Round-2 GPU call Z (1 GPU): the bench line and the reference arm of the final tree (shared workload config, level-(i) leg,
# dual_step_check), smoke.
O=gpurun_out/r02z
mkdir -p $O
python __graft_entry__.py smoke > $O/smoke_final.log 2>&1; echo "smoke exit $?" >> $O/smoke_final.log; tail -2 $O/smoke_final.log
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python bench.py --steps 20 --warmup 3 > $O/bench_final.json 2> $O/bench_final.err; echo "bench exit $?" >> $O/bench_final.err
tail -1 $O/bench_final.err; head -c 400 $O/bench_final.json
