#!/bin/bash
# gpurun --gpus N --timeout 900 -- "bash scripts/gpu_r2_ngpu.sh N": sharded parity tests (NCCL, graphed NCCL, NVLink peer memory) up to R = N and the N-rank bench
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n$N.txt
SEL=${2:-}   # optional pytest -k expression (e.g. "8- or 4-": only the 4- and 8-rank cases)
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q ${SEL:+-k "$SEL"} > gpurun_out/pytest_gpu_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_n$N.log
tail -12 gpurun_out/pytest_gpu_n$N.log
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
nvidia-smi nvlink -s -i 0 >> gpurun_out/topo_n$N.txt 2>&1
timeout 120 python scripts/p2p_bw_probe.py >> gpurun_out/topo_n$N.txt 2>&1; tail -1 gpurun_out/topo_n$N.txt
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_before_n$N.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r02_n$N.log 2>&1; echo "bench$N rc=$?"
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_after_n$N.txt 2>&1
grep -v '^{' gpurun_out/bench_r02_n$N.log | tail -8 | cut -c1-300
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/bench_r02_n{n}.log") if l.startswith("{")][-1])
    print("N", n, "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "timing", d["timing"], "launches", d["gpu_launches"])
    print(" parity", d["parity"])
    print(" nvlink", d.get("nvlink"))
    print(" phases", {k: round(v, 1) for k, v in d.get("kernels_us", {}).items()})
    print(" extras", {k: ({kk: v.get(kk) for kk in ("value", "ms_per_step", "e2e", "error")}) for k, v in d["extra_configs"].items()})
except Exception as e:
    print("unreadable", e)
PY
