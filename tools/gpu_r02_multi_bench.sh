#!/bin/bash
# bench.py at N ranks only (no tests, no NCCL comparison run): gpurun --gpus N -- bash tools/gpu_r02_multi_bench.sh N
N=${1:-2}
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus $N --steps 20 --warmup 3 --compare-nccl > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench exit $?" >> $O/bench_n$N.err
tail -2 $O/bench_n$N.err
