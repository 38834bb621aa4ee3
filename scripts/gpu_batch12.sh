#!/bin/bash
# final sanity of the bench entry points as the driver runs them
mkdir -p gpurun_out
timeout 300 python bench.py --gpus 1 --steps 200 --warmup 20 > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"
timeout 120 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref_final.log 2>&1; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_final.log") if l.startswith("{")][-1])
print("value", round(d["value"]), "e2e", d["e2e"], "ingest", d["ingest"], "cpu", d["cpu_baseline"]["value"], "clocks", d["clocks"], "roofline", d["roofline"]["frac"], d["roofline"]["traffic"])
r = json.loads([l for l in open("gpurun_out/bench_ref_final.log") if l.startswith("{")][-1])
print("ref", r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["sample"][:120])
PY
tail -3 gpurun_out/bench_final.log | cut -c1-200
