#!/bin/bash
# Round-2 GPU call M (2 GPUs): vectorised / batched-poll peer all-reduce -- its test, the distributed step graph, bench at N=2.
O=gpurun_out/r02m
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 > $O/pytest_multi.log 2>&1; echo "pytest exit $? ($(date -u +%Y-%m-%dT%H:%M:%SZ))" >> $O/pytest_multi.log
tail -4 $O/pytest_multi.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 --compare-nccl > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench N=2 exit $?"
python - <<PY
import json
l=json.loads(open("$O/bench_n2.json").read().strip().splitlines()[-1])
print(l["value"], l["grad_allreduce"], l["multi_gpu_check"]["ok"])
PY
