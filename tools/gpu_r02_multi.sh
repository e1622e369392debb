#!/bin/bash
# Round-2 multi-GPU call (gpurun --gpus N): the two-GPU tests with the final kernels, then bench.py at N ranks (weak scaling
# line + multi_gpu_check + strong scaling of config 5 with the in-kernel peer exchange, NCCL variant beside it).
N=${1:-2}
O=gpurun_out/r02
mkdir -p $O
nvidia-smi topo -m > $O/topo_n$N.txt 2>&1
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 > $O/pytest_multi.log 2>&1; echo "pytest exit $? ($(date -u +%Y-%m-%dT%H:%M:%SZ), $(nvidia-smi -L | wc -l) GPUs)" >> $O/pytest_multi.log
  tail -5 $O/pytest_multi.log
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 3 --compare-nccl > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench exit $?" >> $O/bench_n$N.err
tail -3 $O/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 20 --warmup 3 --exchange nccl --skip-e2e --skip-train-step --skip-cpu > $O/bench_n${N}_nccl.json 2> $O/bench_n${N}_nccl.err; echo "bench exit $?" >> $O/bench_n${N}_nccl.err
python - <<PY
import json
for f in ("$O/bench_n$N.json", "$O/bench_n${N}_nccl.json"):
  try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.4g ms/step %.4f" % (d["value"], d["ms_per_step"]), "multi", d.get("multi_gpu_check"), "strong", d.get("strong_scaling"))
  except Exception as e:
    print(f, "unreadable", e)
PY
