#!/bin/bash
mkdir -p gpurun_out
for r in 1 0; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$r bench.py --gpus 2 --steps 200 --warmup 10 --reserve-sms $r --skip-e2e --skip-cpu --skip-train-step > gpurun_out/bench_n2_r$r.json 2> gpurun_out/bench_n2_r$r.err; echo "n2 reserve=$r exit $?"; tail -2 gpurun_out/bench_n2_r$r.err | cut -c1-200; python -c "
import json; d=json.load(open('gpurun_out/bench_n2_r$r.json')); print('reserve',$r, d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
done
timeout 600 python bench.py --steps 200 --warmup 10 --skip-e2e --skip-cpu --skip-train-step > gpurun_out/bench_n1b.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_n1b.json')); print('n1', d['value'], d['ms_per_step'])"
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2
