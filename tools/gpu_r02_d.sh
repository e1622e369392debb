#!/bin/bash
# Round-2 GPU call D (1 GPU): interleaved A/B of the operand-economy variants (split 3-operand packed ops, ALU slope selection),
# the full GPU suite on the new default, train-step profile, bench line.
O=gpurun_out/r02
mkdir -p $O
python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu_d.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_d.log
tail -6 $O/pytest_gpu_d.log
python tools/kernel_bench.py --libs libikl.so,libikl_nosplit.so,libikl_base.so,libikl_split_alu0.so --repeat 3 \
  --configs 4obs_static,1obs_static,4obs_dynamic,joint_limits --iters 30 > $O/kt_split_variants.jsonl 2> $O/kt_split_variants.err
python tools/profile_train_step.py 20 $O/train_step_profile.txt > /dev/null 2> $O/train_step_profile.err
python bench.py --steps 20 --warmup 3 > $O/bench_d.json 2> $O/bench_d.err; echo "bench exit $?" >> $O/bench_d.err
python tools/profile_target.py 24 planes cfg_joint_limits_and_4obs_static > $O/plain_obs4_stat_planes_v2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lagrangian -s 2 -c 1 -o $O/ncu_obs4_stat_planes_v2 python tools/profile_target.py 24 planes cfg_joint_limits_and_4obs_static > $O/ncu_obs4_stat_planes_v2.log 2>&1
