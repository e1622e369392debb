#!/bin/bash
# Round-2 GPU call F (1 GPU): early barrier probe -- correctness (full suite) and interleaved same-run A/B against the build without it.
O=gpurun_out/r02
mkdir -p $O
python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu_f.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_f.log
tail -4 $O/pytest_gpu_f.log
python tools/kernel_bench.py --libs libikl.so,libikl_noprobe.so --repeat 3 --iters 30 > $O/kt_probe.jsonl 2> $O/kt_probe.err
