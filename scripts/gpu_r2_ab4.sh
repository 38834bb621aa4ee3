#!/bin/bash
# gpurun --timeout 900 -- "bash scripts/gpu_r2_ab4.sh": A/B of the 3xTF32 GEMM pipeline depth (2 stages x 2 CTAs/SM vs 4 stages x 1 CTA/SM) on the local step
mkdir -p gpurun_out
run() {  # name, env, args
  env $2 timeout 300 python bench.py --steps 20 --warmup 5 --no-parity --extra "" $3 > gpurun_out/ab4_$1.log 2>&1
  python - $1 <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/ab4_{n}.log") if l.startswith("{")][-1])
    print(n, "us/step", round(1e3 * d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]))
    print("   phases", {k: round(v, 1) for k, v in d.get("kernels_us", {}).items()})
except Exception as e:
    print(n, "unreadable", e); print(open(f"gpurun_out/ab4_{n}.log").read()[-800:])
PY
}
DEEP="PS_B200_LIB=$PWD/ps_b200/lib/libps_b200_deep.so"
run cfg2_base X=1 "--config cfg2"
run cfg2_deep $DEEP "--config cfg2"
run cfg4_base X=1 "--config cfg4"
run cfg4_deep $DEEP "--config cfg4"
run cfg3_base X=1 "--config cfg3"
run cfg3_deep $DEEP "--config cfg3"
