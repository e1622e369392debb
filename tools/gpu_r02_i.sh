#!/bin/bash
# Round-2 GPU call I (2 GPUs): peer-memory all-reduce of a flat vector, the distributed step graph built on it, the other
# two-GPU tests, and a diagnostic of the captured NCCL all-reduce that hung earlier this round.
O=gpurun_out/r02
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 240 -k "peer_memory_all_reduce" > $O/pytest_multi_i1.log 2>&1; echo "pytest exit $?" >> $O/pytest_multi_i1.log
tail -5 $O/pytest_multi_i1.log | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 240 -k "cuda_graph" > $O/pytest_multi_i2.log 2>&1; echo "pytest exit $?" >> $O/pytest_multi_i2.log
tail -5 $O/pytest_multi_i2.log | cut -c1-300
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 -k "not cuda_graph and not peer_memory_all_reduce" > $O/pytest_multi_i3.log 2>&1; echo "pytest exit $?" >> $O/pytest_multi_i3.log
tail -5 $O/pytest_multi_i3.log | cut -c1-300
for mode in thread_local global; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/debug_graph_dist.py $mode > $O/debug_graph_dist_$mode.log 2>&1
  echo "debug_graph_dist $mode exit $?" | tee -a $O/debug_graph_dist_$mode.log
  grep "rank 0" $O/debug_graph_dist_$mode.log | tail -4
done
