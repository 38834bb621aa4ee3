#!/bin/bash
# gpurun --timeout 900 -- "bash scripts/gpu_r2_ab6.sh": GEMM pipeline depth inside the SHARDED step (R = 1 on one GPU), and PDL along the GEMM chain with the deep pipeline
mkdir -p gpurun_out
run() {  # name, args, env...
  n=$1; a=$2; shift 2
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-parity --no-kernel-times --extra "" $a > gpurun_out/ab6_$n.log 2>&1
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/ab6_{n}.log") if l.startswith("{")][-1])
    print(n, "us/step", round(1e3 * d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print(n, "unreadable", e); print(open(f"gpurun_out/ab6_{n}.log").read()[-800:])
PY
}
for c in cfg4 cfg2 cfg3; do
  run fs_${c}_f2_w2 "--config $c --force-sharded" PS_TC_DEEP=2 PS_TC_DEEP_WGRAD=2
  run fs_${c}_f2_w0 "--config $c --force-sharded" PS_TC_DEEP=2 PS_TC_DEEP_WGRAD=0
  run fs_${c}_f0_w0 "--config $c --force-sharded" PS_TC_DEEP=0 PS_TC_DEEP_WGRAD=0
done
run local_cfg2_pdlgemm "--config cfg2" PS_PDL_GEMM=1
run local_cfg4_pdlgemm "--config cfg4" PS_PDL_GEMM=1
