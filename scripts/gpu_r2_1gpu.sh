#!/bin/bash
# gpurun --timeout 900 -- "bash scripts/gpu_r2_1gpu.sh": round-2 record on one B200 — parity tests, smoke, bench (driver-style 20/5) + reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 60 python scripts/tf32_rounding_probe.py > gpurun_out/tf32_probe.log 2>&1; cat gpurun_out/tf32_probe.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_20_5.log 2>&1; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r02_20_5.log | grep -v '^{' | tail -5
python - <<'PY'
import json
for f in ("bench_r02_20_5",):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.log") if l.startswith("{")][-1])
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "timing", d["timing"], "clocks", d["clocks"])
        print(" roofline", d["roofline"] and (d["roofline"]["kernel"], round(d["roofline"]["frac"], 3)), "tensor", d["roofline_tensor"] and (d["roofline_tensor"]["kernel"], round(d["roofline_tensor"]["frac"], 3)), "tf32 peak", d["tf32_peak_tflops"])
        print(" emb", {k: (round(v["us"], 2), round(v["frac"], 3)) for k, v in d["hbm_kernels"].items() if k.startswith("emb") and isinstance(v, dict)}, "resolve", d["hbm_kernels"].get("emb_resolve_only_us"))
        print(" gemm", {k: round(v["us"], 2) for k, v in d["hbm_kernels"].items() if k.startswith("fc")})
        lg = d.get("roofline_large_batch") or {}
        for k in ("zipf", "uniform"):
            if k in lg:
                print(" large", k, {kk: (round(v["us"], 1), round(v["frac"], 3)) for kk, v in lg[k].items() if isinstance(v, dict)}, lg[k]["emb_resolve_only_us"], lg[k]["unique_keys_per_batch"], lg[k]["ring_working_set_mb"])
        print(" parity", d["parity"])
        print(" extras", d["extra_configs"])
        print(" cfg5", d["cfg5"] and (round(d["cfg5"]["value"]), {k: round(v["us"], 2) for k, v in d["cfg5"]["gemms"].items()}))
        print(" ingest", d["ingest"])
        print(" cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], "phases", d["kernels_us"])
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r02.log 2>&1; echo "ref rc=$?"
grep '^{' gpurun_out/bench_ref_r02.log | tail -1 | cut -c1-300
