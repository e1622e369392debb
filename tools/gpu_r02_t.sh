#!/bin/bash
# Round-2 multi-GPU call T (gpurun --gpus N): two-GPU tests (incl. the epoch-closing launch across ranks), bench.py at N ranks
# with the default exchange (once per dual step, inside the closing launch) and with the per-step in-kernel exchange beside it.
N=${1:-2}
O=gpurun_out/r02t
mkdir -p $O
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 400 > $O/pytest_multi.log 2>&1; echo "pytest exit $? ($(date -u +%Y-%m-%dT%H:%M:%SZ), $(nvidia-smi -L | wc -l) GPUs)" >> $O/pytest_multi.log
  tail -5 $O/pytest_multi.log
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench exit $?" >> $O/bench_n$N.err
tail -3 $O/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 20 --warmup 3 --exchange peer --skip-e2e --skip-train-step --skip-cpu --skip-strong --skip-multi-check > $O/bench_n${N}_perstep.json 2> $O/bench_n${N}_perstep.err; echo "bench exit $?" >> $O/bench_n${N}_perstep.err
python - <<PY
import json
for f in ("$O/bench_n$N.json", "$O/bench_n${N}_perstep.json"):
  try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.4g ms/step %.4f kernel_ms %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"]), "dual", d.get("dual_step_check"), "multi", (d.get("multi_gpu_check") or {}).get("ok"), "strong", d.get("strong_scaling"))
  except Exception as e:
    print(f, "unreadable", e)
PY
