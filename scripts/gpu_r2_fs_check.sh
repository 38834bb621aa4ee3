#!/bin/bash
# gpurun --timeout 900 -- "bash scripts/gpu_r2_fs_check.sh": sharded-step tests that fit one GPU (R = 1 self-exchange) + the phase times,
# with the per-block fence of the producer kernels at device scope (default) and at system scope (the old form)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/pytest_fs_check.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_fs_check.log
echo "== device-scope block fences"; bash scripts/gpu_r2_fs_phases.sh
echo "== system-scope block fences"; PS_P2P_BLOCK_FENCE_SYS=1 bash scripts/gpu_r2_fs_phases.sh
