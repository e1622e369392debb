#!/bin/bash
# Round-2 GPU call X (1 GPU): torch.profiler breakdown of level (i) (unmodified reference trainer over the CUDA drop-ins).
O=gpurun_out/r02x
mkdir -p $O
python tools/profile_dropin_step.py $O/dropin_step_profile.txt > $O/dropin.log 2>&1; echo "exit $?" >> $O/dropin.log
grep -n "batch 2^" $O/dropin_step_profile.txt; tail -3 $O/dropin.log
