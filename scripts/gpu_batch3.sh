#!/bin/bash
# one 1-GPU gpurun call: parity tests, the bench line, the reference arm, the ncu launch list and one --set full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_r01.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_r01.log | cut -c1-900
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r01.log 2>&1; tail -1 gpurun_out/bench_ref_r01.log | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv \
  python bench.py --steps 4 --warmup 3 --cpu-budget 0.2 --large '' --no-kernel-times > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'emb_|wide_|dense_update' -s 40 -c 14 -f -o gpurun_out/prof_r01 \
  python bench.py --steps 4 --warmup 3 --cpu-budget 0.2 --large '' --no-kernel-times > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
