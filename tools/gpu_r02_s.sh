#!/bin/bash
# Round-2 GPU call S (1 GPU): the launch that closes a dual step (ikl_lagrangian_*_dual_step): GPU suite, smoke, bench line.
O=gpurun_out/r02s
mkdir -p $O
python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu.log 2>&1; echo "pytest exit $? ($(date -u +%Y-%m-%dT%H:%M:%SZ))" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -2 $O/smoke.log
python bench.py --steps 20 --warmup 3 --skip-cpu --skip-train-step > $O/bench.json 2> $O/bench.err; echo "bench exit $?" >> $O/bench.err
tail -3 $O/bench.err; head -c 1500 $O/bench.json
