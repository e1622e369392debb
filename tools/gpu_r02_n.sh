#!/bin/bash
# Round-2 GPU call N (1 GPU): resident blocks per SM of the light (no-obstacle) staged kernels: cap 2 / 3 / 4 / none.
O=gpurun_out/r02n
mkdir -p $O
for rep in 1 2; do
for cap in 0 2 3 4 5; do
  IKL_GRID_BLOCKS_PER_SM=$cap python tools/kernel_bench.py --configs joint_limits,no_constraints,1obs_static --iters 30 | sed "s/^{/{\"cap\": $cap, /" >> $O/kt_cap.jsonl
done
done
python - <<PY
import json,collections
best=collections.defaultdict(lambda:1e9)
for l in open("$O/kt_cap.jsonl"):
  if not l.startswith("{"): continue
  r=json.loads(l)
  if "mode" not in r: continue
  k=(r["config"][4:],r["layout"],r["mode"],r["cap"]); best[k]=min(best[k],r["ms_best"])
seen=[]
for (c,l,m,_) in best:
  if (c,l,m) in seen: continue
  seen.append((c,l,m))
  print("%-26s %-7s %-13s"%(c,l,m)," ".join("%.4f"%best[(c,l,m,cap)] for cap in (0,2,3,4,5)))
PY
