#!/usr/bin/env python
"""bench.py — Wide&Deep CTR training throughput through libps_b200.so (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2]

A "step" is one train.Trainer step (pullWeights → forward → loss → backward → KVStore.update →
clear) of the WideDeepNN on one synthetic Criteo-libsvm-shaped batch.
  value : samples/s with the batch ring already resident in HBM (device-timed, CUDA events on
          the library's stream, max over ranks)
  e2e   : samples/s through the host-facing C-ABI call (ps_model_submit / ps_model_collect, the
          pipelined form of Trainer.train) with inputs in pinned HOST memory: every step copies
          its E/X/W/Y host→device and reads its loss device→host inside the timed region
  roofline : the dominant HBM-bound kernel's algorithmic bytes / its device time vs the measured
          copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline : the CPU oracle (C++ restatement of the reference's standalone Java path; the JVM
          cannot run in this image) timed on this box's host cores on a bounded sample
--impl reference times that CPU restatement as the reference arm (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from ps_b200.synth import CONFIGS, Synth  # noqa: E402

METRIC = "Wide&Deep CTR training samples/sec"
UNIT = "samples/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region: NVML polled every 2 ms (the timed region of the default run is tens
    of milliseconds — an nvidia-smi subprocess per sample would see it once); nvidia-smi is the fallback when NVML is unavailable."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.sm_max, self.how = index, [], False, None, "nvml"
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv, self.how = None, "nvidia-smi"

    def run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                if nv is not None:
                    sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    self.rows.append((sm, r))
                    time.sleep(0.002)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    t = [x.strip() for x in out.split(",")]
                    r = 0
                    for i, bit in enumerate([0x8, 0x40, 0x20, 0x4]):       # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap (NVML bit values)
                        if t[3 + i].lower().startswith("active"):
                            r |= bit
                    self.sm_max = float(t[1])
                    self.rows.append((float(t[0]), r))
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unsampled"]}
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for _, r in self.rows:
            bits |= r
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.sm_max, "reasons": [n for b, n in names.items() if bits & b],
                "samples": len(self.rows), "how": self.how}


def model_args(cfg, n_gpus):
    return dict(kind=cfg["kind"], F=cfg["F"], D=cfg["D"], Xn=cfg["Xn"], fc=cfg["fc"])


def oracle_model(cfg, seed):
    import oracle_lib as ol
    kind = {"widedeep": ol.KIND_WIDEDEEP, "dnn": ol.KIND_DNN, "fcnn": ol.KIND_FCNN}[cfg["kind"]]
    return ol.OracleModel(kind, cfg["F"], cfg["D"], cfg["Xn"], cfg["fc"], seed, emb_opt=1 if cfg["emb_opt"] == "ftrl" else 0)


def time_oracle(cfg, batches, budget_s, threads):
    """Timed CPU restatement: returns (samples/s, steps run, gemm back-end)."""
    import oracle_lib as ol
    os.environ["OMP_NUM_THREADS"] = str(threads)
    gemm = "openmp-loops"
    ob = ol.openblas_path()
    if ob and ol.lib().pso_set_gemm(2, ob.encode()) == 0:
        gemm = "openblas-0.3.30 (numpy bundled)"
        os.environ["OPENBLAS_NUM_THREADS"] = str(threads)
    else:
        ol.lib().pso_set_gemm(1, None)
    o = oracle_model(cfg, 20261017)
    b = batches[0]
    o.train_step(b.get("E"), b["X"], b.get("W"), b["Y"])          # warm-up: creates keys
    n, t0 = 0, time.perf_counter()
    while True:
        b = batches[(n + 1) % len(batches)]
        o.train_step(b.get("E"), b["X"], b.get("W"), b["Y"])
        n += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or n >= 64:
            break
    ol.lib().pso_set_gemm(0, None)
    return n * cfg["B"] / dt, n, gemm, dt


def time_oracle_replicas(cfg, batches, budget_s, replicas):
    """`replicas` independent copies of the standalone Trainer step (thread = 1 each, single-threaded sgemm), one per host thread:
    an UPPER bound for the reference's Trainer with thread = replicas, whose replicas share one synchronized KVStore
    (KVStore.java:136,192,240) and serialise on it.  ctypes releases the GIL inside the C call, so the steps run in parallel."""
    import oracle_lib as ol
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    gemm = "openmp-loops"
    ob = ol.openblas_path()
    if ob and ol.lib().pso_set_gemm(2, ob.encode()) == 0:
        gemm = "openblas-0.3.30 (numpy bundled), 1 thread per replica"
    else:
        ol.lib().pso_set_gemm(1, None)
    models = [oracle_model(cfg, 20261017 + r) for r in range(replicas)]
    for r, o in enumerate(models):                                   # warm-up: creates keys
        b = batches[r % len(batches)]
        o.train_step(b.get("E"), b["X"], b.get("W"), b["Y"])
    counts = [0] * replicas
    t0 = time.perf_counter()

    def work(r):
        o, n = models[r], 0
        while time.perf_counter() - t0 < budget_s and n < 64:
            b = batches[(r + n + 1) % len(batches)]
            o.train_step(b.get("E"), b["X"], b.get("W"), b["Y"])
            n += 1
        counts[r] = n
    ths = [threading.Thread(target=work, args=(r,)) for r in range(replicas)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    ol.lib().pso_set_gemm(0, None)
    return sum(counts) * cfg["B"] / dt, sum(counts), gemm, dt


def time_ingest(ps, cfg, batch, threads):
    """libsvm text -> the step's staging layout through the native reader: one synthetic batch written as text in the
    reference's own line format (label, F "idx:1" columns, Xn "idx:value" columns), read back for about a second on the host cores."""
    import tempfile
    F, Xn, B = cfg["F"], cfg["Xn"], cfg["B"]
    if not F:
        return None
    lines = []
    for n in range(min(B, 4096)):
        cols = ["%d" % int(batch["Y"][n])] + ["%d:1" % int(batch["E"][n, j]) for j in range(F)] + ["%d:%.2f" % (33895 + x, batch["X"][n, x]) for x in range(Xn)]
        lines.append(" ".join(cols))
    with tempfile.NamedTemporaryFile("w", suffix=".libsvm", delete=False) as f:
        f.write("\n".join(lines) + "\n")
        path = f.name
    try:
        r = ps.LibsvmReader(path, F=F, Xn=Xn, batch=len(lines), threads=threads)
        bufs = dict(E=np.empty((len(lines), F), np.int64), W=np.empty((len(lines), F), np.int64), X=np.empty((len(lines), Xn), np.float32), Y=np.empty(len(lines), np.float32))
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 1.0:
            b = r.next(bufs)
            if b is None:
                r.reset()
                continue
            n += len(b["Y"])
        dt = time.perf_counter() - t0
        ok = bool(np.array_equal(bufs["E"][: len(lines)], batch["E"][: len(lines)]) and np.array_equal(bufs["X"][: len(lines)], batch["X"][: len(lines)]))
        r.close()
        return {"lines_per_s": n / dt, "threads": threads, "bytes_per_line": os.path.getsize(path) / len(lines), "roundtrip_exact": ok,
                "sample": f"{n} lines of {len(lines)}-line synthetic libsvm text in {dt:.2f}s through ps_reader_next"}
    finally:
        os.unlink(path)


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    syn = Synth(F=cfg["F"], Xn=cfg["Xn"], V=cfg["V"], dist=args.dist, seed=20261017 + 2)
    batches = [syn.batch(cfg["B"]) for _ in range(4)]
    threads = os.cpu_count() or 1
    per_step_budget = 4.0
    total = args.steps + args.warmup
    sps, n, gemm, dt = time_oracle_replicas(cfg, batches, min(60.0, per_step_budget * total), threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * cfg["B"] / sps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_name(args, cfg), "global_batch": cfg["B"]},
        "cpu_baseline": {"value": sps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n} Trainer steps of batch {cfg['B']} in {dt:.1f}s over {threads} independent replicas (one per host thread, thread=1 each): "
                                   f"an upper bound for the reference's Trainer with thread={threads}, which shares one synchronized KVStore; C++ restatement of the "
                                   f"reference's standalone Java path (JVM/jblas unavailable in this image); sgemm={gemm}"},
        "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def large_batch_roofline(ps, ctx, dev, hbm_peak, name, seed):
    """The embedding kernels at a batch large enough to be bandwidth- rather than launch-latency-bound (cfg4 shapes:
    B=16384, D=64, 10 M keys): same kernels, same measurement (replayed in a CUDA graph, CUDA events on the library's stream)."""
    import torch
    cfg = dict(CONFIGS[name])
    B, F, D, Xn, V = cfg["B"], cfg["F"], cfg["D"], cfg["Xn"], cfg["V"]
    model = ps.Model(ctx, cfg["kind"], F, D, Xn, cfg["fc"], emb_capacity=2 * V + (1 << 16), max_batch=B)
    syn = Synth(F=F, Xn=Xn, V=V, dist="zipf", seed=seed)
    ring = [syn.batch(B) for _ in range(8)]
    dev_ring = [{k: torch.from_numpy(np.ascontiguousarray(v)).cuda(dev) for k, v in b.items()} for b in ring]
    torch.cuda.synchronize()

    def p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None
    for _ in range(2):                       # create the ring's keys, reach the steady state of the table
        for d in dev_ring:
            model.train_step_dev(p(d.get("E")), p(d["X"]), p(d.get("W")), p(d["Y"]), B)
    model.read_loss()
    kt = model.kernel_times([d["E"].data_ptr() for d in dev_ring], B, reps=32)
    L = B * F
    uniq = float(np.mean([len(np.unique(b["E"])) for b in ring]))
    out = {"workload": f"{name} shapes: B={B} F={F} D={D} vocab={V} zipf, {uniq:.0f} unique keys/batch", "peak": hbm_peak, "unit": "GB/s"}
    for k, bytes_ in {"emb_gather": L * (8 + 8 * D), "emb_scatter_update": L * (8 + 4 * D) + uniq * 24 * D}.items():
        us = kt[k] + (kt["emb_probe"] if k == "emb_gather" else 0.0)
        out[k] = {"us": us, "alg_bytes": bytes_, "gbs": bytes_ / max(us, 1e-3) / 1e3, "frac": bytes_ / max(us, 1e-3) / 1e3 / hbm_peak}
    out["emb_probe_us"] = kt["emb_probe"]
    model.close()
    return out


def workload_name(args, cfg):
    return (f"{args.config}: {cfg['kind']} synthetic Criteo-libsvm, F={cfg['F']} Xn={cfg['Xn']} D={cfg['D']} vocab={cfg['V']} "
            f"fc={cfg['fc']} batch={cfg['B']}/GPU emb_opt={cfg['emb_opt']} keys={args.dist}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--dist", default="zipf")
    # FcLayer arithmetic: tf32x3 = tcgen05 tensor cores with error-compensated operand split (fp32-grade results, the
    # default: the reference computes in fp32); tf32 = plain TF32 tensor cores; fp32 = FFMA exact mode
    ap.add_argument("--precision", default=os.environ.get("PS_FC_PRECISION", "tf32x3"))
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--ring", type=int, default=16)
    ap.add_argument("--slack", type=float, default=2.0)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--force-sharded", action="store_true", help="use the sharded step even with one rank (profiling)")
    ap.add_argument("--large", default="cfg4", help="shapes for the large-batch embedding roofline ('' = skip)")
    ap.add_argument("--no-kernel-times", action="store_true", help="skip the per-kernel replays (ncu runs: only the step's own launches)")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        import __graft_entry__ as g
        if rank == 0:
            g.build()
            run_reference(args, cfg, rank)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    if local_rank == 0:
        g.build()
    if world > 1 or args.force_sharded:
        if world == 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    from ps_b200 import binding as ps

    torch.cuda.set_device(local_rank)
    B, F, D, Xn = cfg["B"], cfg["F"], cfg["D"], cfg["Xn"]
    ctx = ps.Context(local_rank, seed=20261017)
    ctx.set_fc_precision({"fp32": ps.PS_FC_FP32, "tf32": ps.PS_FC_TF32, "tf32x3": ps.PS_FC_TF32X3}[args.precision])
    cap = int(min(2 ** 31 - 1, max(1 << 16, (2 * cfg["V"]) // world + (1 << 16)))) if cfg["V"] else 1024
    upd = ps.UpdaterSpec.ftrl() if cfg["emb_opt"] == "ftrl" else None
    model = ps.Model(ctx, cfg["kind"], F, D, Xn, cfg["fc"], emb_capacity=cap, emb_updater=upd, max_batch=B)

    syn = Synth(F=F, Xn=Xn, V=cfg["V"], dist=args.dist, seed=20261017 + 2 + 1000 * rank, n_classes=10 if cfg["kind"] == "fcnn" else 0)
    ring = [syn.batch(B) for _ in range(args.ring)]
    stream = torch.cuda.ExternalStream(ctx.stream(), device=local_rank)

    def to_dev(b):
        d = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda(local_rank) for k, v in b.items()}
        return d
    dev_ring = [to_dev(b) for b in ring]
    torch.cuda.synchronize()

    def ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    # N > 1: the embedding table is sharded by key hash over the ranks and the R ranks perform ONE
    # Trainer step on the concatenated batch (ps_b200/sharded.py); per-GPU batch fixed => weak scaling
    trainer = None
    if world > 1 or args.force_sharded:
        from ps_b200.sharded import GpuOps, GraphedShardedTrainer, P2PShardedTrainer
        if args.exchange == "p2p":
            # every exchange is a store into the consumer's HBM over NVLink by the kernel that produced the data
            trainer = P2PShardedTrainer(ps, ctx, model, rank, world, B, F, slack=args.slack, device=local_rank)
        else:
            # the whole sharded step (local kernels + NCCL collectives, fixed-capacity buckets) is one CUDA graph per rank
            trainer = GraphedShardedTrainer(GpuOps(ps, ctx, model, local_rank), rank, world, B, F, cfg["kind"] == "widedeep", slack=args.slack)

    def dev_step(i):
        d = dev_ring[i % len(dev_ring)]
        if trainer is not None:
            return trainer.step(d.get("E"), d["X"], d.get("W"), d["Y"])
        model.train_step_dev(ptr(d.get("E")), ptr(d["X"]), ptr(d.get("W")), ptr(d["Y"]), B)

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----
    for i in range(args.warmup):
        dev_step(i)
    loss = model.read_loss()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        dev_step(args.warmup + i)
    e1.record(stream)
    loss = model.read_loss()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    if trainer is not None:
        if hasattr(trainer, "launches_per_step"):
            launches = args.steps * trainer.launches_per_step               # torch graph replays: counted at the eager warm-up step
    if world > 1:
        t = torch.tensor([ms_dev], device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev = float(t.item())

    # ---- end-to-end through the host-facing call ("e2e") ----
    pinned = []
    for b in ring:
        pb = {}
        for k, v in b.items():
            pa = ps.PinnedArray(v.shape, v.dtype)
            pa.array[...] = v
            pb[k] = pa
        pinned.append(pb)

    def hp(pb, k):
        return pb[k].ptr if k in pb else None
    h2d = sum(pa.nbytes for pa in pinned[0].values())

    def host_loop(n, start):
        if trainer is not None and args.exchange == "p2p":
            # peer-memory sharded step through its host-facing call: ps_model_p2p_submit copies this rank's slice from pinned host
            # memory and enqueues the step's graph, ps_model_collect returns the global loss of the oldest of 2 steps in flight
            for i in range(n):
                pb = pinned[(start + i) % len(pinned)]
                model.p2p_submit_ptrs(hp(pb, "E"), hp(pb, "X"), hp(pb, "W"), hp(pb, "Y"), B)
                if i >= 1:
                    model.collect()
            return model.collect()
        if trainer is not None:      # NCCL sharded step: stage this rank's slice from pinned host memory, then the step, then its loss
            last = None
            for i in range(n):
                pb = pinned[(start + i) % len(pinned)]
                with torch.cuda.stream(stream):
                    d = {k: torch.from_numpy(pa.array).to(f"cuda:{local_rank}", non_blocking=True) for k, pa in pb.items()}
                trainer.step(d.get("E"), d["X"], d.get("W"), d["Y"])
                last = model.read_loss()                       # the step's result is read back every step
            return last
        for i in range(n):
            pb = pinned[(start + i) % len(pinned)]
            model.submit_ptrs(hp(pb, "E"), hp(pb, "X"), hp(pb, "W"), hp(pb, "Y"), B)
            if i >= 1:
                model.collect()
        return model.collect()
    host_loop(max(3, args.warmup), 0)
    barrier()
    t0 = time.perf_counter()
    loss_e2e = host_loop(args.steps, args.warmup)
    ctx.synchronize()
    t1 = time.perf_counter()
    ms_e2e = (t1 - t0) * 1e3
    if world > 1:
        t = torch.tensor([ms_e2e], device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    clocks = sampler.summary()
    if trainer is not None:
        trainer.check()

    # ---- per-kernel device times and the roofline of the dominant HBM-bound kernel ----
    acc = {}
    reps = 20 if (world == 1 and trainer is None and not args.no_kernel_times) else 0
    if reps:
        model.profile(True)
    for i in range(reps):
        dev_step(i)
        model.read_loss()
        for k, v in model.phase_times().items():
            acc.setdefault(k, []).append(v)
    model.profile(False)
    phase_us = {k: 1e3 * float(np.median(v)) for k, v in acc.items()}
    L = B * F
    uniq = float(np.mean([len(np.unique(b["E"] + (np.arange(F, dtype=np.int64) << 44)[None, :])) for b in ring])) if F else 0.0
    hbm_peak, peak_src = peaks()
    kernels, roofline = {}, None
    if F and world == 1 and trainer is None and not args.no_kernel_times:
        # each embedding kernel replayed 64x inside a CUDA graph over the batch ring, CUDA events on the library's stream
        kt = model.kernel_times([d["E"].data_ptr() for d in dev_ring], B, reps=64)
        alg = {"emb_gather": L * (8 + 8 * D), "emb_scatter_update": L * (8 + 4 * D) + uniq * 24 * D}
        for k, bytes_ in alg.items():
            us = kt[k] + (kt["emb_probe"] if k == "emb_gather" else 0.0)     # the gather's key resolution is the probe kernel
            kernels[k] = {"us": us, "alg_bytes": bytes_, "gbs": bytes_ / max(us, 1e-3) / 1e3}
        kernels["emb_probe"] = {"us": kt["emb_probe"]}
        gt = model.gemm_times(B, reps=64)
        flops = {}
        dims = [F * D + Xn] + list(cfg["fc"])
        for l in range(len(cfg["fc"])):
            for n in ("forward", "dgrad", "wgrad"):
                us = gt[f"fc{l}.{n}"]
                kernels[f"fc{l}.{n}"] = {"us": us, "tflops": 2.0 * B * dims[l] * dims[l + 1] / max(us, 1e-3) / 1e6}
        dom = max(alg, key=lambda k: kernels[k]["us"])
        traffic = None                      # dram__bytes_read+write per launch from the committed ncu --set full capture
        tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tp) and args.config == "cfg2":
            tk = json.load(open(tp))["kernels"]      # emb_scatter_update = the scatter launch + the update launch; gather includes its probe
            parts = {"emb_gather": ["emb_probe_kernel", "emb_gather_kernel"], "emb_scatter_update": ["emb_scatter_kernel", "emb_update_kernel"]}[dom]
            traffic = sum(tk[p]["dram_bytes_per_launch"] for p in parts) if all(p in tk for p in parts) else None
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": kernels[dom]["gbs"] / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                    "how": "algorithmic bytes (SURVEY 8d) / mean device time of the kernel replayed 64x in a CUDA graph over the batch ring "
                           "(CUDA events on the library's stream); emb_gather includes its probe kernel"}

    large = None
    if roofline is not None and args.large:
        large = large_batch_roofline(ps, ctx, local_rank, hbm_peak, args.large, 20261017 + 4)

    if rank == 0:
        cpu = None
        if world == 1:
            sps, n, gemm, dt = time_oracle(cfg, ring[:4], args.cpu_budget, 1)
            cpu = {"value": sps, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"{n} Trainer steps of batch {B} in {dt:.1f}s, thread=1 (CTR.java:72); C++ restatement of the reference's standalone "
                             f"Java path (no JVM in this image); sgemm={gemm} single-threaded"}
        ingest = None
        if world == 1 and trainer is None:
            try:
                ingest = time_ingest(ps, cfg, ring[0], min(16, os.cpu_count() or 1))
            except Exception as e:        # the reader is host-side plumbing: never fail the bench line over it
                ingest = {"error": str(e)}
        total = B * world * args.steps
        line = {
            "metric": METRIC, "value": total / (ms_dev / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "tf32": "tf32", "tf32x3": "f32 via 3xTF32 tcgen05"}[args.precision], "data": "synthetic",
            "config": {"workload": workload_name(args, cfg), "global_batch": B * world,
                       "parallelism": (f"key-hash sharded embedding table over {world} GPUs, exchange={args.exchange} "
                                       f"({'NVLink peer-memory stores, no collective calls' if args.exchange == 'p2p' else 'NCCL all-to-all'}) "
                                       "+ data-parallel dense") if world > 1 else "single",
                       "l2": "embedding table + optimiser state (%.0f MB) exceeds the 126 MB L2; a ring of %d distinct batches; no flush" % (
                           cap * (16 + 12 * D) / 1e6, len(ring))},
            "e2e": {"value": total / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32,
                    "api": "ps_model_submit/ps_model_collect (2 steps in flight)" if trainer is None else
                           "ps_model_p2p_submit/ps_model_collect (2 steps in flight)" if args.exchange == "p2p" else
                           "pinned host batch -> device (async copy on the step's stream) -> sharded step -> ps_model_read_loss, every step"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_large_batch": large, "kernels_us": phase_us, "hbm_kernels": kernels,
            "cpu_baseline": cpu, "ingest": ingest, "loss": loss, "loss_e2e": loss_e2e, "unique_keys_per_batch": uniq,
        }
        print(json.dumps(line), flush=True)
    sys.stdout.flush()
    if world > 1 or trainer is not None:
        # captured NCCL graphs + communicator teardown order is fragile: everything is measured and printed, leave hard
        ctx.synchronize()
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)
    for pb in pinned:
        for pa in pb.values():
            pa.free()
    model.close()
    ctx.close()


if __name__ == "__main__":
    main()
