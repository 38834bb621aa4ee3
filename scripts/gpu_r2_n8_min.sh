#!/bin/bash
# gpurun --gpus 8 --timeout 60 -- "bash scripts/gpu_r2_n8_min.sh": cfg2 on 8 GPUs, default settings, no parity leg, no extra configs (fits a one-minute slot)
mkdir -p gpurun_out
PYTHONFAULTHANDLER=1 timeout 50 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29677 bench.py --gpus 8 --steps 20 --warmup 5 --no-parity --extra "" > gpurun_out/bench_n8_min.log 2>&1; echo "rc=$?"
grep '^{' gpurun_out/bench_n8_min.log | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('N 8 us/step', round(1e3*d['ms_per_step'],1), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
except Exception as e: print('no line', e)"
grep -n "Fatal Python error\|Segmentation\|File \"" gpurun_out/bench_n8_min.log | head -20
