#!/bin/bash
# gpurun --gpus 2 --timeout 900 -- "bash scripts/gpu_round_2gpu.sh": sharded parity tests (NCCL, graphed NCCL, NVLink peer memory) and the 2-rank bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus2.txt
timeout 420 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
tail -12 gpurun_out/pytest_gpu2.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/bench_2gpu.log 2>&1; echo "bench2 rc=$?"
grep '^{' gpurun_out/bench_2gpu.log | tail -1 | cut -c1-700
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 200 --warmup 20 --exchange nccl > gpurun_out/bench_2gpu_nccl.log 2>&1; echo "bench2 nccl rc=$?"
grep '^{' gpurun_out/bench_2gpu_nccl.log | tail -1 | cut -c1-400
tail -5 gpurun_out/bench_2gpu.log | cut -c1-300
