#!/bin/bash
# gpurun --timeout 600 -- "bash scripts/gpu_r2_ab7.sh": 128-column tiles for GEMMs whose 64-column grid is between one and two waves (fc0 dgrad at cfg2: 224 CTAs)
mkdir -p gpurun_out
run() {  # name, args, env...
  n=$1; a=$2; shift 2
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-parity --no-kernel-times --extra "" $a > gpurun_out/ab7_$n.log 2>&1
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/ab7_{n}.log") if l.startswith("{")][-1])
    print(n, "us/step", round(1e3 * d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print(n, "unreadable", e); print(open(f"gpurun_out/ab7_{n}.log").read()[-800:])
PY
}
for c in cfg2 cfg3 cfg4; do
  run ${c}_rule0 "--config $c" PS_TC_WIDE_RULE=0
  run ${c}_rule1 "--config $c" PS_TC_WIDE_RULE=1
done
