#!/bin/bash
# Round-2 GPU call H (1 GPU): gradient-only kernels -- parity tests on the default and the IKL_PREMUL builds, then a same-run
# A/B of the variants (default / premul / 3 resident blocks for the gradient-only static kernels), list cycled twice.
O=gpurun_out/r02
mkdir -p $O
L=$PWD/lagrange-dual-learning_b200/lib
python -m pytest tests/test_gpu_parity.py tests/test_gpu_trainer.py -m gpu -q --timeout 900 -x \
  -k "gradient_only or ties_through or unit_upstream or training_steps_use or graph" > $O/pytest_h.log 2>&1; echo "pytest exit $?" >> $O/pytest_h.log
tail -5 $O/pytest_h.log
for v in premul gonly3; do
  IKL_LIB=$L/libikl_$v.so python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 \
    -k "gradient_only or ties_through or randomised or process_batch" > $O/pytest_h_$v.log 2>&1; echo "pytest exit $?" >> $O/pytest_h_$v.log
  tail -3 $O/pytest_h_$v.log
done
python tools/kernel_bench.py --libs libikl.so,libikl_premul.so,libikl_gonly3.so,libikl_premul_gonly3.so --repeat 2 \
  --configs 4obs_dynamic,4obs_static,1obs_static,joint_limits,no_constraints --iters 30 > $O/kt_gradonly.jsonl 2> $O/kt_gradonly.err
echo "kernel_bench exit $?"
