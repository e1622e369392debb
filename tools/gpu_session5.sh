#!/bin/bash
# two-GPU session: NCCL trainer test + bench at N=2 (and N=1 for the ratio)
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi.log; tail -5 gpurun_out/pytest_multi.log
timeout 600 python bench.py --steps 100 --warmup 5 --skip-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/bench_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 exit $?"; tail -3 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
