#!/bin/bash
# one gpurun call: parity tests, step timeline, bench variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python scripts/trace_step.py --out gpurun_out/trace_x3.json > gpurun_out/trace_x3.log 2>&1; echo "trace rc=$?"
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_b1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_b1.log | cut -c1-600
PS_STREAM_PRIO=0 timeout 300 python bench.py --steps 200 --warmup 20 --cpu-budget 1 --large '' > gpurun_out/bench_b1_noprio.log 2>&1; tail -1 gpurun_out/bench_b1_noprio.log | cut -c1-300
timeout 300 python bench.py --steps 200 --warmup 20 --cpu-budget 1 --large '' --precision tf32 > gpurun_out/bench_b1_tf32.log 2>&1; tail -1 gpurun_out/bench_b1_tf32.log | cut -c1-300
