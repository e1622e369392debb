#!/bin/bash
# Round-2 GPU call Y (1 GPU): function-level drop-ins with the stride-0 gradient and the lighter Python wrappers: GPU suite,
# then the level-(i) profile again.
O=gpurun_out/r02y
mkdir -p $O
python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu.log 2>&1; echo "pytest exit $? ($(date -u +%Y-%m-%dT%H:%M:%SZ))" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
python tools/profile_dropin_step.py $O/dropin_step_profile.txt > $O/dropin.log 2>&1; echo "exit $?" >> $O/dropin.log
grep -n "batch 2^" $O/dropin_step_profile.txt; tail -2 $O/dropin.log
