#!/bin/bash
# gpurun --timeout 900 -- "bash scripts/gpu_r2_fs_check.sh": sharded-step tests that fit one GPU (R = 1 self-exchange) + the phase times,
# with the owner/dense update deferred to the head of the next step (default) and inside the step (PS_P2P_DEFER=0)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/pytest_fs_check.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_fs_check.log
PS_P2P_DEFER=0 timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_fs_check_nodefer.log 2>&1; echo "pytest(no defer) rc=$?"
tail -3 gpurun_out/pytest_fs_check_nodefer.log
echo "== deferred"; bash scripts/gpu_r2_fs_phases.sh
echo "== in the step"; PS_P2P_DEFER=0 bash scripts/gpu_r2_fs_phases.sh
