#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/bench_2gpu.log 2>&1; echo "bench2 rc=$?"
grep '^{' gpurun_out/bench_2gpu.log | tail -1 | cut -c1-1500
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_2gpu.log 2>&1; echo "ref2 rc=$?"
grep '^{' gpurun_out/bench_ref_2gpu.log | tail -1 | cut -c1-300
