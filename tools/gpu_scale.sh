#!/bin/bash
# scaling check on an N-GPU box: bench at 1, 2, 4, 8 ranks (torchrun for N > 1), one JSON line each
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  if [ $n -gt $NG ]; then break; fi
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "N=$n exit $?"; tail -2 gpurun_out/scale_n$n.err | cut -c1-200
  python - <<PY
import json
l=[x for x in open('gpurun_out/scale_n$n.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('N=$n value %.4g ms/step %.4f e2e %.4g train %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['train_step'] and round(d['train_step']['steps_per_s'],2)))
PY
done
