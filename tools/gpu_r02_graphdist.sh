#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
IKL_GRAPH_DISTRIBUTED=1 timeout 240 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 200 -k "cuda_graph" > $O/pytest_multi_graph.log 2>&1; echo "pytest exit $?" >> $O/pytest_multi_graph.log
tail -30 $O/pytest_multi_graph.log | cut -c1-300
