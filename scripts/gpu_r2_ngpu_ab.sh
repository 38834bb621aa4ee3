#!/bin/bash
# gpurun --gpus N --timeout 900 -- "bash scripts/gpu_r2_ngpu_ab.sh N": the N-rank bench with the update deferred (default) and in the step
N=${1:-2}
mkdir -p gpurun_out
for d in 1 0; do
  PS_P2P_DEFER=$d timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$d bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_ab_defer${d}_n$N.log 2>&1; echo "bench defer=$d rc=$?"
  python - $N $d <<'PY'
import json, sys
n, dd = sys.argv[1], sys.argv[2]
try:
    d = json.loads([l for l in open(f"gpurun_out/bench_ab_defer{dd}_n{n}.log") if l.startswith("{")][-1])
    print("N", n, "defer", dd, "us/step", round(1e3 * d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]), "parity", d["parity"]["ok"])
    print(" phases", {k: round(v, 1) for k, v in d.get("kernels_us", {}).items()})
    print(" extras", {k: round(1e3 * v.get("ms_per_step", 0), 1) for k, v in d["extra_configs"].items()})
except Exception as e:
    print("unreadable", e); print(open(f"gpurun_out/bench_ab_defer{dd}_n{n}.log").read()[-1500:])
PY
done
