#!/bin/bash
# gpurun --timeout 1100 -- "bash scripts/gpu_round_1gpu.sh": the round record on one B200 — parity tests, bench + reference arm, launch list, full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_r01.log 2>&1; echo "bench rc=$?"
timeout 300 python bench.py --steps 200 --warmup 20 --cpu-budget 0.5 --large '' --force-sharded > gpurun_out/bench_p2p_1rank.log 2>&1; echo "bench(force-sharded) rc=$?"
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r01.log 2>&1; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv \
  python bench.py --steps 4 --warmup 3 --cpu-budget 0.2 --large '' --no-kernel-times --ring 4 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'emb_|wide_|dense_update|gemm_tf32' -s 120 -c 26 -f -o gpurun_out/prof_r01 \
  python bench.py --steps 4 --warmup 3 --cpu-budget 0.2 --large '' --no-kernel-times --ring 4 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:emb_ -s 64 -c 4 -f -o gpurun_out/prof_large \
  python scripts/large_batch_steps.py > gpurun_out/ncu_large.log 2>&1; echo "ncu large rc=$?"
python - <<'PY'
import json
for f in ("bench_r01", "bench_p2p_1rank"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.log") if l.startswith("{")][-1])
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "clocks", d["clocks"],
              "roofline", d["roofline"] and (d["roofline"]["kernel"], round(d["roofline"]["frac"], 3)),
              {k: round(v["us"], 2) for k, v in d["hbm_kernels"].items() if k.startswith("emb")},
              "large", d.get("roofline_large_batch") and {k: (round(v["us"], 1), round(v["frac"], 3)) for k, v in d["roofline_large_batch"].items() if isinstance(v, dict)})
    except Exception as e:
        print(f, "unreadable", e)
PY
