#!/bin/bash
# Round-2 GPU call R (1 GPU): final reduction with all loads in one round trip -- parity (sums are bit-compared between
# layouts / modes in the suite), then the graph-captured step on a 2^21-sample shard and on 2^24.
O=gpurun_out/r02r
mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 -x > $O/pytest_r.log 2>&1; echo "pytest exit $?" >> $O/pytest_r.log
tail -3 $O/pytest_r.log
for rep in 1 2 3; do
  for lg in 21 24; do
    python bench.py --steps 50 --warmup 5 --skip-e2e --skip-train-step --skip-cpu --skip-context --strong-total-log2 $lg 2> /dev/null | \
      python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'strong_log2': $lg, 'ms_per_step': l['ms_per_step'], 'graph_ms': l['strong_scaling']['kernel_only_ms']}))" >> $O/reduce_ab.jsonl
  done
done
cat $O/reduce_ab.jsonl
