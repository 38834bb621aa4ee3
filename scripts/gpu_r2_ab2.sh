#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[1].split("/")[-1], "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "parity", d["parity"] and d["parity"]["ok"])
    print("  gemm", {k: round(v["us"], 1) for k, v in d["hbm_kernels"].items() if k.startswith("fc")})
    print("  cfg5", d["cfg5"] and (round(d["cfg5"]["value"]), {k: round(v["us"], 1) for k, v in d["cfg5"]["gemms"].items()}))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
timeout 400 python bench.py --steps 20 --warmup 5 --extra '' --large '' > gpurun_out/bench_base.log 2>&1; summ gpurun_out/bench_base.log
PS_GEMM_NARROW=1 timeout 400 python bench.py --steps 20 --warmup 5 --extra '' --large '' > gpurun_out/bench_narrow.log 2>&1; summ gpurun_out/bench_narrow.log
