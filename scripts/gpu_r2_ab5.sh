#!/bin/bash
# gpurun --timeout 900 -- "bash scripts/gpu_r2_ab5.sh": 3xTF32 GEMM pipeline depth per launch: PS_TC_DEEP (forward/dgrad) x PS_TC_DEEP_WGRAD
mkdir -p gpurun_out
run() {  # name, env..., args
  env $2 $3 timeout 300 python bench.py --steps 20 --warmup 5 --no-parity --no-kernel-times --extra "" --config $4 > gpurun_out/ab5_$1.log 2>&1
  python - $1 <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/ab5_{n}.log") if l.startswith("{")][-1])
    print(n, "us/step", round(1e3 * d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print(n, "unreadable", e); print(open(f"gpurun_out/ab5_{n}.log").read()[-800:])
PY
}
for c in cfg2 cfg3 cfg4; do
  for v in "0 0" "2 0" "2 2" "2 1" "1 1"; do
    set -- $v
    run ${c}_f$1_w$2 PS_TC_DEEP=$1 PS_TC_DEEP_WGRAD=$2 $c
  done
done
