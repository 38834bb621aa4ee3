#!/bin/bash
# gpurun --timeout 1200 -- "bash scripts/gpu_r2_iter.sh": iteration check — tests, bench with embedding rooflines, sharded-on-one-rank timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 5 --extra '' > gpurun_out/bench_iter.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_iter.log") if l.startswith("{")][-1])
    print("value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "timing", d["timing"])
    print(" emb", {k: (round(v["us"], 2), round(v["frac"], 3)) for k, v in d["hbm_kernels"].items() if k.startswith("emb") and isinstance(v, dict)}, "resolve", d["hbm_kernels"].get("emb_resolve_only_us"))
    lg = d.get("roofline_large_batch") or {}
    for k in ("zipf", "uniform"):
        if k in lg:
            print(" large", k, {kk: (round(v["us"], 1), round(v["frac"], 3)) for kk, v in lg[k].items() if isinstance(v, dict)}, lg[k]["emb_resolve_only_us"])
    print(" parity", d["parity"] and d["parity"]["ok"], "ingest", d["ingest"] and d["ingest"].get("text_to_train", {}).get("lines_per_s"))
except Exception as e:
    print("unreadable", e)
PY
tail -3 gpurun_out/bench_iter.log | grep -v '^{' | cut -c1-300
for c in cfg2 cfg4; do
  timeout 300 python bench.py --config $c --steps 20 --warmup 5 --force-sharded --no-kernel-times --no-parity --ring 8 > gpurun_out/bench_fs_$c.log 2>&1; echo "force-sharded $c rc=$?"
  grep '^{' gpurun_out/bench_fs_$c.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  fs', d['config']['workload'][:12], 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))"
  PS_DEBUG_NO_UCNT=1 timeout 300 python bench.py --config $c --steps 20 --warmup 5 --force-sharded --no-kernel-times --no-parity --ring 8 > gpurun_out/bench_fs_noucnt_$c.log 2>&1
  grep '^{' gpurun_out/bench_fs_noucnt_$c.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  fs-noucnt', d['config']['workload'][:12], 'value', round(d['value']), 'ms', round(d['ms_per_step'],4))"
  timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-kernel-times --no-parity --ring 8 > gpurun_out/bench_local_$c.log 2>&1
  grep '^{' gpurun_out/bench_local_$c.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  local', d['config']['workload'][:12], 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_fs_cfg4.csv \
  python bench.py --config cfg4 --steps 3 --warmup 2 --reps 2 --force-sharded --no-kernel-times --no-parity --ring 4 > gpurun_out/ncu_l3.log 2>&1; echo "ncu fs cfg4 rc=$?"
python scripts/ncu_summary.py launches gpurun_out/launches_fs_cfg4.csv gpurun_out/launches_fs_cfg4.md; head -24 gpurun_out/launches_fs_cfg4.md | cut -c1-140
