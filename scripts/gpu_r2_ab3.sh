#!/bin/bash
# gpurun --timeout 900 -- "bash scripts/gpu_r2_ab3.sh": A/B of PS_PDL_EXCHANGE on one GPU (local step and the sharded step with R = 1)
mkdir -p gpurun_out
PS_PDL_EXCHANGE=1 timeout 300 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/pytest_ab3.log 2>&1; echo "pytest(pdl exchange) rc=$?"; tail -3 gpurun_out/pytest_ab3.log
run() {  # name, env, args
  env $2 timeout 300 python bench.py --steps 20 --warmup 5 --no-parity --no-kernel-times --extra "" $3 > gpurun_out/ab3_$1.log 2>&1
  python - $1 <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/ab3_{n}.log") if l.startswith("{")][-1])
    print(n, "us/step", round(1e3 * d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print(n, "unreadable", e); print(open(f"gpurun_out/ab3_{n}.log").read()[-800:])
PY
}
run local_cfg2_off PS_PDL_EXCHANGE=0 "--config cfg2"
run local_cfg2_on PS_PDL_EXCHANGE=1 "--config cfg2"
run fs_cfg2_off PS_PDL_EXCHANGE=0 "--config cfg2 --force-sharded"
run fs_cfg2_on PS_PDL_EXCHANGE=1 "--config cfg2 --force-sharded"
run local_cfg4_off PS_PDL_EXCHANGE=0 "--config cfg4"
run local_cfg4_on PS_PDL_EXCHANGE=1 "--config cfg4"
run fs_cfg4_off PS_PDL_EXCHANGE=0 "--config cfg4 --force-sharded"
run fs_cfg4_on PS_PDL_EXCHANGE=1 "--config cfg4 --force-sharded"
