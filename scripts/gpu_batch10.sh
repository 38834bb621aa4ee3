#!/bin/bash
# quick 1-GPU check of the pipelined peer-memory e2e path (R = 1) + smoke
mkdir -p gpurun_out
timeout 200 python bench.py --steps 200 --warmup 20 --cpu-budget 0.5 --large '' --force-sharded > gpurun_out/bench_p2p_1rank.log 2>&1; echo "bench(force-sharded) rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python - <<'PY'
import json
for f in ("bench_p2p_1rank",):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.log") if l.startswith("{")][-1])
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", d["e2e"], "loss", d["loss"], d["loss_e2e"])
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -5 gpurun_out/bench_p2p_1rank.log | cut -c1-300
