#!/bin/bash
# gpurun --timeout 1500 -- "bash scripts/gpu_r2_ab.sh": tests + bench + A/B switches (grouped wgrad, TMA for every row, register-path scatter / update)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[1].split("/")[-1], "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]))
    print("  emb", {k: (round(v["us"], 2), round(v["frac"], 3)) for k, v in d["hbm_kernels"].items() if k.startswith("emb") and isinstance(v, dict)})
    lg = d.get("roofline_large_batch") or {}
    for k in ("zipf", "uniform"):
        if k in lg:
            print("  large", k, {kk: (round(v["us"], 1), round(v["frac"], 3)) for kk, v in lg[k].items() if isinstance(v, dict)}, lg[k]["emb_resolve_only_us"])
    print("  gemm", {k: round(v["us"], 1) for k, v in d["hbm_kernels"].items() if k.startswith("fc")}, "phases", {k: round(v, 1) for k, v in d["kernels_us"].items()})
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
timeout 400 python bench.py --steps 20 --warmup 5 --extra '' > gpurun_out/bench_iter.log 2>&1; echo "bench rc=$?"; summ gpurun_out/bench_iter.log
PS_GROUP_WGRAD=0 timeout 400 python bench.py --steps 20 --warmup 5 --extra '' --large '' --no-parity > gpurun_out/bench_nogroup.log 2>&1; summ gpurun_out/bench_nogroup.log
PS_HOT_SHARE=1 timeout 400 python bench.py --steps 20 --warmup 5 --extra '' --no-parity > gpurun_out/bench_alltma.log 2>&1; summ gpurun_out/bench_alltma.log
PS_SCATTER_SLAB=0 PS_UPDATE_SLAB=0 timeout 400 python bench.py --steps 20 --warmup 5 --extra '' --no-parity > gpurun_out/bench_noslab.log 2>&1; summ gpurun_out/bench_noslab.log
PS_UPDATE_SLAB=0 timeout 400 python bench.py --steps 20 --warmup 5 --extra '' --no-parity > gpurun_out/bench_noupdslab.log 2>&1; summ gpurun_out/bench_noupdslab.log
