#!/bin/bash
# Round-2 multi-GPU call U (gpurun --gpus N): bench.py at N ranks with the default exchange (once per dual step, inside the
# closing launch); multi_gpu_check and strong scaling included, host-buffer / train-step / CPU legs skipped.
N=${1:-8}
O=gpurun_out/r02u
mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 3 --skip-e2e --skip-train-step --skip-cpu > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench exit $?" >> $O/bench_n$N.err
tail -2 $O/bench_n$N.err
python - <<PY
import json
d = json.loads(open("$O/bench_n$N.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f kernel_ms %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"]), "dual", d.get("dual_step_check"), "multi", d.get("multi_gpu_check"), "strong", d.get("strong_scaling"))
PY
