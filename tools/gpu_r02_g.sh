#!/bin/bash
# Round-2 GPU call G (1 GPU): direct-load kernels (no staging) on the static configs, for comparison with the staged ones.
O=gpurun_out/r02
mkdir -p $O
python tools/kernel_bench.py --configs 4obs_static,1obs_static,joint_limits,no_constraints --iters 30 > $O/kt_staged_now.jsonl 2> $O/kt_staged_now.err
IKL_PLANES_TMA=0 IKL_ROWS_TMA=0 python tools/kernel_bench.py --configs 4obs_static,1obs_static,joint_limits,no_constraints --iters 30 > $O/kt_direct.jsonl 2> $O/kt_direct.err
python - <<PY
import json
a={ (r['config'],r['layout'],r['mode']):r for r in map(json.loads,[l for l in open("$O/kt_staged_now.jsonl") if l.startswith('{')])}
b={ (r['config'],r['layout'],r['mode']):r for r in map(json.loads,[l for l in open("$O/kt_direct.jsonl") if l.startswith('{')])}
for k in a:
  print(k, a[k]['ms_best'], b[k]['ms_best'], b[k]['kernel'][:40])
PY
