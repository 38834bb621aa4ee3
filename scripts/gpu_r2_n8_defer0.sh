#!/bin/bash
# gpurun --gpus 8 --timeout 600 -- "bash scripts/gpu_r2_n8_defer0.sh": cfg2 on 8 GPUs with the owner/dense update inside the step (PS_P2P_DEFER=0), no extras
mkdir -p gpurun_out
PS_P2P_DEFER=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 --steps 20 --warmup 5 --no-parity --extra "" > gpurun_out/bench_n8_defer0.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n8_defer0.log") if l.startswith("{")][-1])
    print("N 8 defer 0 us/step", round(1e3 * d["ms_per_step"], 1), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]))
    print(" phases", {k: round(v, 1) for k, v in d.get("kernels_us", {}).items()})
except Exception as e:
    print("unreadable", e); print(open("gpurun_out/bench_n8_defer0.log").read()[-1500:])
PY
