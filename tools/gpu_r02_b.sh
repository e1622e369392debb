#!/bin/bash
# Round-2 GPU call B (1 GPU): parity of the single-sample static kernels, same-run A/B of the occupancy / packing variants on
# the static-obstacle configs, and the measured parity report at 2^24.
O=gpurun_out/r02
mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_trainer.py -m gpu -q --timeout 900 > $O/pytest_gpu_b.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_b.log
tail -15 $O/pytest_gpu_b.log
python tools/kernel_bench.py --libs libikl.so,libikl_pair2.so,libikl_pair3.so,libikl_single2.so,libikl_single4.so \
  --configs 4obs_static,1obs_static,joint_limits,no_constraints --iters 30 > $O/kt_static_variants.jsonl 2> $O/kt_static_variants.err
python tools/parity_report.py --out $O/parity_report.json > $O/parity_report.log 2>&1; echo "parity exit $?" >> $O/parity_report.log
tail -3 $O/parity_report.log
