#!/bin/bash
# Round-2 GPU call O (1 GPU): programmatic dependent launch (IKL_PDL=1) -- parity, then A/B of the back-to-back bench step and of
# the graph-captured steps on a 2^21-sample shard (the strong-scaling launch of one of eight ranks) and on 2^24.
O=gpurun_out/r02o
mkdir -p $O
IKL_PDL=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trainer.py -m gpu -q --timeout 900 -x > $O/pytest_pdl.log 2>&1; echo "pytest exit $?" >> $O/pytest_pdl.log
tail -3 $O/pytest_pdl.log
for rep in 1 2 3; do
for pdl in 0 1; do
  for lg in 21 24; do
    IKL_PDL=$pdl python bench.py --steps 50 --warmup 5 --skip-e2e --skip-train-step --skip-cpu --skip-context --strong-total-log2 $lg 2> /dev/null | \
      python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'pdl': $pdl, 'strong_log2': $lg, 'ms_per_step': l['ms_per_step'], 'graph_ms': l['strong_scaling']['kernel_only_ms']}))" >> $O/pdl_ab.jsonl
  done
done
done
cat $O/pdl_ab.jsonl
