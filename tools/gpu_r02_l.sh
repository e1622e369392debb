#!/bin/bash
# Round-2 GPU call L (N GPUs, N = $1): the driver's bench command at N ranks with the final kernels.
N=${1:-2}
O=gpurun_out/r02l
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 --compare-nccl > $O/bench_n$N.json 2> $O/bench_n$N.err
echo "bench N=$N exit $?"
tail -c 600 $O/bench_n$N.err
