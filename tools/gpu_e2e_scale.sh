#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" > gpurun_out/lscpu.txt
for bind in yes no; do
for n in 4 8; do
  extra=""; [ $bind = no ] && extra="--no-numa-bind"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 20 --warmup 3 --skip-cpu --skip-train-step --skip-context $extra > gpurun_out/e2e_n${n}_$bind.json 2> gpurun_out/e2e_n${n}_$bind.err
  python - <<PY
import json
l=[x for x in open('gpurun_out/e2e_n${n}_$bind.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('N=$n bind=$bind e2e %.4g  (%s)' % (d['e2e']['value'], d['e2e'].get('numa_bind')))
PY
done; done
cat gpurun_out/lscpu.txt; head -12 gpurun_out/topo.txt
