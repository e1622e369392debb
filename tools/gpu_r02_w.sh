#!/bin/bash
# Round-2 GPU call W (1 GPU): bench line with the shared workload config and the level-(i) timing leg (the unmodified reference
# trainer over the function-level CUDA drop-ins); ncu --set full of the benchmark kernel on a 2^21-sample shard (the strong-
# scaling regime at N = 8: where the ~7.5 us that do not shrink with the shard go).
O=gpurun_out/r02w
mkdir -p $O
python bench.py --steps 20 --warmup 3 --skip-train-step > $O/bench.json 2> $O/bench.err; echo "bench exit $?" >> $O/bench.err
tail -2 $O/bench.err
python -c "
import json
d = json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['context'].get('reference_trainer_over_dropins'), d['context'].get('reference_eager_cuda'))
"
python tools/profile_target.py 21 planes cfg_joint_limits_and_4obs_dynamic fwdbwd > $O/plain_2p21.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lagrangian -s 2 -c 1 -o $O/ncu_dyn4_planes_2p21 python tools/profile_target.py 21 planes cfg_joint_limits_and_4obs_dynamic fwdbwd > $O/ncu_2p21.log 2>&1
python tools/ncu_summarise.py $O/ncu_dyn4_planes_2p21.ncu-rep $O/ncu_dyn4_planes_2p21 > /dev/null 2>&1
rm -f $O/ncu_dyn4_planes_2p21.ncu-rep
ls -la $O
