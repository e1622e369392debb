#!/bin/bash
# Round-2 closing GPU call (1 GPU, the last minutes of the budget): GPU suite + smoke of the final tree, with the wall time of
# each leg.  Bounded by its own timeouts.
O=gpurun_out/r02final
mkdir -p $O
date -u +%FT%TZ > $O/when.txt
( time timeout 150 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
( time timeout 60 python __graft_entry__.py smoke ) > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
tail -3 $O/smoke.log
