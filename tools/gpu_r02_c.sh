#!/bin/bash
# Round-2 GPU call C (1 GPU): full GPU suite on the default build, smoke, the ALU-select variants' parity, same-run A/B of
# ALU-select / 128-thread-block variants, J=2 / runtime-length kernel table rows, and a short bench line (compact e2e inputs).
O=gpurun_out/r02
mkdir -p $O
./tools/microbench/fma_forms > $O/microbench_fma_forms.txt 2>&1
python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu_c.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_c.log
tail -6 $O/pytest_gpu_c.log
python __graft_entry__.py smoke > $O/smoke_c.log 2>&1; echo "smoke exit $?" >> $O/smoke_c.log; tail -2 $O/smoke_c.log
IKL_LIB=$PWD/lagrange-dual-learning_b200/lib/libikl_alu2.so python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 -k "not two_links" > $O/pytest_gpu_c_alu2.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_c_alu2.log
tail -6 $O/pytest_gpu_c_alu2.log
python tools/kernel_bench.py --libs libikl.so,libikl_alu1.so,libikl_alu2.so,libikl_t128s4.so,libikl_t128s5.so,libikl_t128s5alu2.so \
  --configs 4obs_static,1obs_static,4obs_dynamic --iters 30 > $O/kt_alu_variants.jsonl 2> $O/kt_alu_variants.err
python tools/kernel_bench.py --num-links 2 --iters 30 --configs 4obs_dynamic,4obs_static,no_constraints > $O/kt_j2.jsonl 2> $O/kt_j2.err
python tools/kernel_bench.py --link-lengths 0.7,1.3,0.45 --iters 30 --configs 4obs_dynamic,4obs_static > $O/kt_len.jsonl 2> $O/kt_len.err
python tools/profile_train_step.py 20 $O/train_step_profile.txt > /dev/null 2> $O/train_step_profile.err
python bench.py --steps 20 --warmup 3 --skip-train-step > $O/bench_c.json 2> $O/bench_c.err; echo "bench exit $?" >> $O/bench_c.err
