#!/bin/bash
# gpurun --timeout 1800 -- "bash scripts/gpu_r2_final_1gpu.sh": the round-2 record on one B200 — parity tests, smoke, bench (driver-style and long),
# reference arm, launch lists, ncu --set full captures (cfg2 step kernels; embedding kernels at cfg4 shapes, zipf and uniform)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_20_5.log 2>&1; echo "bench 20/5 rc=$?"
timeout 400 python bench.py --steps 200 --warmup 20 --extra '' > gpurun_out/bench_r02_200_20.log 2>&1; echo "bench 200/20 rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r02.log 2>&1; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("bench_r02_20_5", "bench_r02_200_20", "bench_ref_r02"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.log") if l.startswith("{")][-1])
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), d.get("timing"), d.get("clocks"))
        if d.get("roofline"):
            print("  roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), "tensor", round(d["roofline_tensor"]["frac"], 3), "tf32", round(d["tf32_peak_tflops"]))
            lg = d.get("roofline_large_batch") or {}
            for k in ("zipf", "uniform"):
                if k in lg:
                    print("  large", k, {kk: (round(v["us"], 1), round(v["frac"], 3)) for kk, v in lg[k].items() if isinstance(v, dict)})
            print("  parity", d["parity"] and d["parity"]["ok"], "extras", {k: (round(v.get("value", 0)), round(v.get("ms_per_step", 0), 4)) for k, v in (d.get("extra_configs") or {}).items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02_cfg2.csv \
  python bench.py --steps 3 --warmup 2 --reps 2 --no-kernel-times --no-parity --ring 4 > gpurun_out/ncu_l1.log 2>&1; echo "ncu launches cfg2 rc=$?"
timeout 500 ncu --set full --clock-control none -k regex:'emb_|gemm_tf32|dense_update|wide_' -s 150 -c 22 -f -o gpurun_out/prof_r02_cfg2 \
  python bench.py --steps 3 --warmup 2 --reps 2 --no-kernel-times --no-parity --ring 4 > gpurun_out/ncu_f1.log 2>&1; echo "ncu full cfg2 rc=$?"
timeout 500 ncu --set full --clock-control none -k regex:'emb_lookup|emb_scatter_slab|emb_update_slab' -s 48 -c 6 -f -o gpurun_out/prof_r02_large \
  python scripts/large_batch_steps.py cfg4 4 > gpurun_out/ncu_large.log 2>&1; echo "ncu large rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r02_fs_cfg2.csv \
  python bench.py --steps 3 --warmup 2 --reps 2 --force-sharded --no-kernel-times --no-parity --ring 4 > gpurun_out/ncu_l2.log 2>&1; echo "ncu launches sharded cfg2 rc=$?"
for f in r02_cfg2 r02_fs_cfg2; do python scripts/ncu_summary.py launches gpurun_out/launches_$f.csv gpurun_out/launches_$f.md; done
# the reports themselves exceed what gpurun brings back (64 MiB): summarise them here, keep the text
python scripts/ncu_summary.py full gpurun_out/prof_r02_cfg2.ncu-rep gpurun_out/ncu_full_cfg2.md
python scripts/ncu_summary.py traffic gpurun_out/prof_r02_cfg2.ncu-rep gpurun_out/traffic_cfg2.json
python scripts/ncu_summary.py full gpurun_out/prof_r02_large.ncu-rep gpurun_out/ncu_full_large.md
for r in prof_r02_cfg2 prof_r02_large; do
  ncu -i gpurun_out/$r.ncu-rep --page details --csv 2>/dev/null | grep -i -E "stall|Issue Slots|Eligible|No Eligible|One or More" | cut -c1-400 | head -400 > gpurun_out/${r}_details.csv
done
ls -la gpurun_out/*.ncu-rep; rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
