#!/bin/bash
# gpurun --timeout 600 -- "bash scripts/gpu_r2_fs_phases.sh": the sharded (peer-memory) step on ONE GPU (R = 1, self-exchange) with per-phase
# main-stream times, beside the local step: what the exchange pipeline itself costs, without NVLink
mkdir -p gpurun_out
for c in cfg2 cfg4; do
  timeout 300 python bench.py --config $c --steps 20 --warmup 5 --force-sharded --no-parity --extra "" > gpurun_out/fs_phases_$c.log 2>&1; echo "fs $c rc=$?"
  python - $c <<'PY'
import json, sys
c = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/fs_phases_{c}.log") if l.startswith("{")][-1])
    print(c, "value", round(d["value"]), "us/step", round(1e3 * d["ms_per_step"], 1), "launches", d["gpu_launches"])
    print("  phases", {k: round(v, 1) for k, v in d["kernels_us"].items()}, "sum", round(sum(d["kernels_us"].values()), 1))
except Exception as e:
    print("unreadable", e)
    print(open(f"gpurun_out/fs_phases_{c}.log").read()[-1500:])
PY
done
