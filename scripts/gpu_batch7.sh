#!/bin/bash
# 1-GPU call: parity tests, bench, ncu of the embedding kernels at cfg4 shapes
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_b7.log 2>&1; echo "bench rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:emb_ -s 64 -c 4 -f -o gpurun_out/prof_large \
  python scripts/large_batch_steps.py > gpurun_out/ncu_large.log 2>&1; echo "ncu large rc=$?"
python - <<'PY'
import json
for f in ("bench_b7",):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.log") if l.startswith("{")][-1])
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3),
              {k: round(v["us"], 2) for k, v in d["hbm_kernels"].items() if k.startswith("emb")},
              "large", {k: (round(v["us"], 1), round(v["frac"], 3)) for k, v in d["roofline_large_batch"].items() if isinstance(v, dict)}, d["roofline_large_batch"].get("emb_probe_us"),
              "phases", {k: round(v, 1) for k, v in d["kernels_us"].items() if k.startswith("emb")})
    except Exception as e:
        print(f, "unreadable", e)
PY
