#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 120 python scripts/tf32_rounding_probe.py > gpurun_out/tf32_rounding.log 2>&1; cat gpurun_out/tf32_rounding.log
timeout 300 python scripts/trace_step.py --out gpurun_out/trace_x3.json > gpurun_out/trace_x3.log 2>&1; echo "trace rc=$?"
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_b2.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_b2.log | cut -c1-300
