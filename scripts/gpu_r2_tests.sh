#!/bin/bash
# gpurun --timeout 900 -- "bash scripts/gpu_r2_tests.sh": the whole GPU test suite + smoke + one driver-style bench line on one GPU
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_20_5.log 2>&1; echo "bench 20/5 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_r02_20_5.log") if l.startswith("{")][-1])
    print("value", round(d["value"]), "us", round(1e3 * d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]), "parity", d["parity"]["ok"],
          "extras", {k: round(1e3 * v.get("ms_per_step", 0), 1) for k, v in d["extra_configs"].items()}, "clocks", d["clocks"])
except Exception as e:
    print("unreadable", e); print(open("gpurun_out/bench_r02_20_5.log").read()[-1500:])
PY
