#!/bin/bash
# gpurun --gpus 2 --timeout 500 -- "bash scripts/gpu_r2_n2_segv.sh": does the crash seen at 8 GPUs with --no-parity --extra "" (profiles/r02_notes.md) show at two? (it did not)
mkdir -p gpurun_out
for d in 0 2; do
  PS_P2P_DEFER=$d PYTHONFAULTHANDLER=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2966$d bench.py --gpus 2 --steps 20 --warmup 5 --no-parity --extra "" > gpurun_out/segv_defer$d.log 2>&1; echo "defer=$d rc=$?"
  grep -c '^{' gpurun_out/segv_defer$d.log
  grep -n "Fatal Python error\|File \"\|Segmentation\|Current thread\|Thread 0x" gpurun_out/segv_defer$d.log | head -30
done
