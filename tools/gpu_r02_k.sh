#!/bin/bash
# Round-2 GPU call K (1 GPU): near-tie tracking by three-input minima + 32-bit tile indices -- parity, then same-run A/B against
# the previous build (libikl_prev.so), list cycled three times.
O=gpurun_out/r02k
mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 -x > $O/pytest_k.log 2>&1; echo "pytest exit $?" >> $O/pytest_k.log
tail -4 $O/pytest_k.log
python tools/kernel_bench.py --libs libikl_prev.so,libikl.so --repeat 3 \
  --configs 4obs_dynamic,4obs_static,1obs_static,joint_limits,no_constraints --iters 30 > $O/kt_min3.jsonl 2> $O/kt_min3.err
echo "kernel_bench exit $?"
