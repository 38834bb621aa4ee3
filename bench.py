#!/usr/bin/env python
"""bench.py — Wide&Deep CTR training throughput through libps_b200.so (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2]

A "step" is one train.Trainer step (pullWeights → forward → loss → backward → KVStore.update →
clear) of the WideDeepNN on one synthetic Criteo-libsvm-shaped batch.
  value : samples/s with the batch ring already resident in HBM.  Device-timed (CUDA events on the library's stream), max over
          ranks; the K-step loop is repeated (>= 25 times, >= 1 s in total) and the MEDIAN repetition is reported.  Before any
          timing every batch of the ring has been stepped once, so no graph capture and no first-touch insert is timed.
  e2e   : samples/s through the host-facing C-ABI call (ps_model_submit / ps_model_collect, the pipelined form of
          Trainer.train) with inputs in pinned HOST memory: every step copies its E/X/W/Y host→device and reads its loss
          device→host inside the timed region; same repetition / median protocol, wall clock around synchronised regions
  roofline : the dominant HBM-bound kernel's algorithmic bytes / its device time vs the measured copy bandwidth in
          MEASURED_PEAKS.json; roofline_tensor: the dominant FcLayer GEMM vs a TF32 cuBLAS ceiling measured in this run
  parity : a fresh model of the same shape, stepped on a global batch split over the N ranks, against the CPU oracle's step on
          the concatenated batch (loss, dense weights, >= 200 embedding rows fetched from whichever rank owns them)
  extra_configs : BASELINE configs 3 (100 M keys, Ftrl, D = 32) and 4 (DNN, 10 M keys, D = 64) at this N, same protocol
  cpu_baseline : the CPU oracle (C++ restatement of the reference's standalone Java path; the JVM cannot run in this image)
          timed on this box's host cores on a bounded sample
--impl reference times that CPU restatement as the reference arm (rank 0 only).
"""
import os
import sys

# The oracle's sgemm is the OpenBLAS that numpy bundles; numpy initialises it with one thread per core at import.  Both the
# single-thread baseline and the one-replica-per-core reference arm want ONE BLAS thread per call: pin it before numpy loads
# (and again, explicitly, on the dlopen'ed handle — see pin_blas()).
os.environ["OMP_NUM_THREADS"] = "1"
os.environ["OPENBLAS_NUM_THREADS"] = "1"

import argparse  # noqa: E402
import ctypes as C  # noqa: E402
import json  # noqa: E402
import math  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from ps_b200.synth import CONFIGS, Synth  # noqa: E402

METRIC = "Wide&Deep CTR training samples/sec"
UNIT = "samples/s"
L2_BYTES = 126e6
DEPTH = 4                            # steps the host keeps in flight through ps_model_submit / collect (the library stages 4)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region: NVML polled every 5 ms; nvidia-smi is the fallback when NVML is unavailable."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index, threaded=True):
        """threaded = False (multi-rank runs): no second Python thread in the process — poll() is called by the main thread while the GPU
        works on the steps it has just enqueued (see main(): a collection started from the sampler thread ended two 8-GPU runs)."""
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.sm_max, self.how = index, [], False, None, "nvml"
        self.threaded, self.started = threaded, False
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

    def begin(self):
        if self.threaded:
            self.started = True
            self.start()
        else:
            self.how += ", polled by the main thread between the enqueue and the synchronisation of a repetition"

    def poll(self):
        """One sample, from the calling thread (inline mode only; NVML: ~20 us)."""
        nv = self.nv
        if self.threaded or nv is None:
            return
        try:
            sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            self.rows.append((sm, r))
        except Exception:
            pass

    def run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                if nv is not None:
                    sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    self.rows.append((sm, r))
                    time.sleep(0.005)
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
        if self.started:
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


NVLINK_WHY = []                      # why the NVLink counters could not be read (reported in the JSON line instead of a silent null)


def nvlink_kib(index):
    """Cumulative NVLink payload KiB (tx, rx) of one GPU over all its links; None when neither NVML nor nvidia-smi can say (the reason
    goes to NVLINK_WHY).  Tries the NVML field counters (all-links scope, then per link), then `nvidia-smi nvlink -gt d`."""
    why = []
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        out = []
        for fid in (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX):
            v = pynvml.nvmlDeviceGetFieldValues(h, [(fid, 0xFFFFFFFF)])[0]           # scope UINT_MAX = all links
            if v.nvmlReturn == 0:
                out.append(int(v.value.ullVal))
                continue
            per_link = pynvml.nvmlDeviceGetFieldValues(h, [(fid, l) for l in range(18)])   # NV18: sum the links
            good = [int(x.value.ullVal) for x in per_link if x.nvmlReturn == 0]
            if not good:
                why.append("NVML field %d: nvmlReturn %d (all links), %s (per link)" % (fid, v.nvmlReturn, sorted({int(x.nvmlReturn) for x in per_link})))
                out = None
                break
            out.append(sum(good))
        if out is not None:
            return tuple(out)
    except Exception as e:
        why.append("NVML: %s: %s" % (type(e).__name__, e))
    try:
        import re
        import subprocess
        txt = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True, timeout=20).stdout
        tx = [int(m) for m in re.findall(r"Data Tx:\s*(\d+)\s*KiB", txt)]
        rx = [int(m) for m in re.findall(r"Data Rx:\s*(\d+)\s*KiB", txt)]
        if tx and rx:
            return (sum(tx), sum(rx))
        why.append("nvidia-smi nvlink -gt d: no counters in %r" % txt.strip()[:160])
    except Exception as e:
        why.append("nvidia-smi nvlink: %s: %s" % (type(e).__name__, e))
    NVLINK_WHY[:] = why
    return None


# ----------------------------------------------------------------------------------------------- CPU oracle legs
def oracle_model(cfg, seed):
    import oracle_lib as ol
    kind = {"widedeep": ol.KIND_WIDEDEEP, "dnn": ol.KIND_DNN, "fcnn": ol.KIND_FCNN}[cfg["kind"]]
    return ol.OracleModel(kind, cfg["F"], cfg["D"], cfg["Xn"], cfg["fc"], seed, emb_opt=1 if cfg["emb_opt"] == "ftrl" else 0)


def pin_blas():
    """Selects OpenBLAS sgemm for the oracle and forces ONE thread per call on that very library handle.  Returns (name, threads in force)."""
    import oracle_lib as ol
    ob = ol.openblas_path()
    if ob and ol.lib().pso_set_gemm(2, ob.encode()) == 0:
        n = ol.lib().pso_set_blas_threads(1)
        return "openblas-0.3.30 (numpy bundled)", n
    ol.lib().pso_set_gemm(1, None)
    return "openmp-loops", 1


def time_oracle_replicas(cfg, batches, budget_s, replicas, max_steps=64):
    """`replicas` independent copies of the standalone Trainer step (thread = 1 each, single-threaded sgemm), one per host thread:
    an UPPER bound for the reference's Trainer with thread = replicas, whose replicas share one synchronized KVStore
    (KVStore.java:136,192,240) and serialise on it.  ctypes releases the GIL inside the C call, so the steps run in parallel."""
    import oracle_lib as ol
    gemm, blas_threads = pin_blas()
    models = [oracle_model(cfg, 20261017 + r) for r in range(replicas)]
    for r, o in enumerate(models):                                   # warm-up: creates keys
        b = batches[r % len(batches)]
        o.train_step(b.get("E"), b["X"], b.get("W"), b["Y"])
    counts = [0] * replicas
    t0 = time.perf_counter()

    def work(r):
        o, n = models[r], 0
        while time.perf_counter() - t0 < budget_s and n < max_steps:
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
    return sum(counts) * cfg["B"] / dt, sum(counts), f"{gemm}, {blas_threads} BLAS thread per call", dt


def workload_name(args, cfg, name=None):
    return (f"{name or args.config}: {cfg['kind']} synthetic Criteo-libsvm, F={cfg['F']} Xn={cfg['Xn']} D={cfg['D']} vocab={cfg['V']} "
            f"fc={cfg['fc']} batch={cfg['B']}/GPU emb_opt={cfg['emb_opt']} keys={args.dist}")


def run_reference(args, cfg):
    syn = Synth(F=cfg["F"], Xn=cfg["Xn"], V=cfg["V"], dist=args.dist, seed=20261017 + 2)
    batches = [syn.batch(cfg["B"]) for _ in range(4)]
    threads = os.cpu_count() or 1
    total = args.steps + args.warmup
    single, n1, gemm, dt1 = time_oracle_replicas(cfg, batches, 4.0, 1)
    sps, n, gemm, dt = time_oracle_replicas(cfg, batches, min(45.0, 3.0 * total), threads)
    speedup = sps / single
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * cfg["B"] / sps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_name(args, cfg), "global_batch": cfg["B"] * args.gpus},
        "cpu_baseline": {"value": sps, "unit": UNIT, "cores": threads, "kind": "port",
                         "single_thread_value": single, "replica_speedup": speedup,
                         "replicas_scale": bool(speedup >= 0.5 * threads),     # 1-thread replicas on idle cores must scale: the round-1 arm did not (BLAS oversubscription)
                         "sample": f"{n} Trainer steps of batch {cfg['B']} in {dt:.1f}s over {threads} independent replicas (one per host thread, thread=1 each; "
                                   f"one replica alone: {single:.0f} samples/s, x{speedup:.1f} with {threads}): an upper bound for the reference's Trainer with "
                                   f"thread={threads}, which shares one synchronized KVStore; C++ restatement of the reference's standalone Java path "
                                   f"(JVM/jblas unavailable in this image); sgemm={gemm}"},
        "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- the GPU arm
class Env:
    """Process-wide handles: torch / distributed / the ctypes binding / one library context per rank."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import __graft_entry__ as g
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.sharded = self.world > 1 or args.force_sharded
        if self.local_rank == 0:
            g.build()
        if self.sharded:
            if self.world == 1:
                os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
                os.environ.setdefault("MASTER_PORT", "29533")
                dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", self.local_rank))
            else:
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            dist.barrier()
        from ps_b200 import binding as ps
        self.ps = ps
        torch.cuda.set_device(self.local_rank)
        self.ctx = ps.Context(self.local_rank, seed=20261017)
        self.ctx.set_fc_precision({"fp32": ps.PS_FC_FP32, "tf32": ps.PS_FC_TF32, "tf32x3": ps.PS_FC_TF32X3}[args.precision])
        self.stream = torch.cuda.ExternalStream(self.ctx.stream(), device=self.local_rank)
        self.sampler = None                  # the clock sampler while the headline workload is timed (polled inline in multi-rank runs)

    def barrier(self):
        self.ctx.synchronize()
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        if self.world == 1:
            return list(values)
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=f"cuda:{self.local_rank}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def bcast_int(self, v):
        if self.world == 1:
            return int(v)
        t = self.torch.tensor([int(v)], dtype=self.torch.int64, device=f"cuda:{self.local_rank}")
        self.dist.broadcast(t, 0)
        return int(t.item())


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Workload:
    """One model + its ring of synthetic batches (device-resident and pinned-host copies) + the sharded trainer when N > 1."""

    def __init__(self, env, name, cfg, ring, batch=None, capacity=None, seed_off=0):
        ps, torch = env.ps, env.torch
        self.env, self.name, self.cfg = env, name, cfg
        self.B = B = batch or cfg["B"]
        F, D, Xn = cfg["F"], cfg["D"], cfg["Xn"]
        world = env.world
        self.cap = capacity or (int(min(2 ** 31 - 1, max(1 << 16, (2 * cfg["V"]) // world + (1 << 16)))) if cfg["V"] else 1024)
        upd = ps.UpdaterSpec.ftrl() if cfg["emb_opt"] == "ftrl" else None
        self.model = ps.Model(env.ctx, cfg["kind"], F, D, Xn, cfg["fc"], emb_capacity=self.cap, emb_updater=upd, max_batch=B)
        syn = Synth(F=F, Xn=Xn, V=cfg["V"], dist=env.args.dist, seed=20261017 + 2 + seed_off + 1000 * env.rank, n_classes=10 if cfg["kind"] == "fcnn" else 0)
        self.ring = [syn.batch(B) for _ in range(ring)]
        dev = env.local_rank
        self.dev_ring = [{k: torch.from_numpy(np.ascontiguousarray(v)).cuda(dev) for k, v in b.items()} for b in self.ring]
        torch.cuda.synchronize()
        self.trainer = None
        if env.sharded:
            from ps_b200.sharded import GpuOps, GraphedShardedTrainer, P2PShardedTrainer
            if env.args.exchange == "p2p":
                self.trainer = P2PShardedTrainer(ps, env.ctx, self.model, env.rank, world, B, F, slack=env.args.slack, device=dev)
            else:
                self.trainer = GraphedShardedTrainer(GpuOps(ps, env.ctx, self.model, dev), env.rank, world, B, F, cfg["kind"] == "widedeep", slack=env.args.slack)
        self.pinned = None

    def dev_step(self, i):
        d = self.dev_ring[i % len(self.dev_ring)]
        if self.trainer is not None:
            return self.trainer.step(d.get("E"), d["X"], d.get("W"), d["Y"])
        self.model.train_step_dev(ptr(d.get("E")), ptr(d["X"]), ptr(d.get("W")), ptr(d["Y"]), self.B)

    def prepare(self):
        """Every batch of the ring once, untimed: the step graph of each ring entry is captured and instantiated, its keys are inserted."""
        for i in range(len(self.dev_ring) + 2):
            self.dev_step(i)
        loss = self.model.read_loss()
        self.env.barrier()
        return loss

    def _reps(self, K, probe_ms, min_reps=25, target_s=1.0, max_reps=1500):
        reps = max(min_reps, int(math.ceil(target_s * 1e3 / max(probe_ms, 1e-3))))
        return self.env.bcast_int(min(reps, max_reps))

    def time_value(self, K, warmup, reps=None):
        env, torch = self.env, self.env.torch
        for i in range(warmup):
            self.dev_step(i)
        self.model.read_loss()
        env.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def one(start):
            env.barrier()
            e0.record(env.stream)
            for i in range(K):
                self.dev_step(start + i)
            e1.record(env.stream)
            if env.sampler is not None:
                env.sampler.poll()                   # the GPU is working on the K steps just enqueued
            env.barrier()
            return e0.elapsed_time(e1)
        probe = env.max_over_ranks([one(warmup)])[0]
        reps = reps or self._reps(K, probe)
        l0 = env.ctx.launch_count()
        nv0 = nvlink_kib(env.local_rank) if env.world > 1 else None
        ms = [one(warmup + (r + 1) * K) for r in range(reps)]
        nv1 = nvlink_kib(env.local_rank) if env.world > 1 else None
        self.nvlink = None
        if nv0 and nv1:                              # NVML's own counters: what actually crossed this GPU's links during the timed steps
            self.nvlink = {"tx_bytes_per_step": 1024.0 * (nv1[0] - nv0[0]) / (reps * K), "rx_bytes_per_step": 1024.0 * (nv1[1] - nv0[1]) / (reps * K)}
        launches = (env.ctx.launch_count() - l0) / reps
        if self.trainer is not None and hasattr(self.trainer, "launches_per_step"):
            launches = K * self.trainer.launches_per_step           # torch graph replays: counted at the eager warm-up step
        ms = sorted(env.max_over_ranks(ms))
        return {"ms": ms[len(ms) // 2], "ms_min": ms[0], "ms_max": ms[-1], "reps": reps, "launches": launches, "loss": self.model.read_loss()}

    def _pin(self):
        if self.pinned is None:
            ps = self.env.ps
            self.pinned = []
            for b in self.ring:                      # the same batches as the device-resident leg: same table rows, same L2 behaviour
                pb = {}
                for k, v in b.items():
                    pa = ps.PinnedArray(v.shape, v.dtype)
                    pa.array[...] = v
                    pb[k] = pa
                self.pinned.append(pb)
        return self.pinned

    def host_loop(self, n, start, mid=None):
        """n steps through the host-facing calls; mid: called once, half way (the inline clock sample of a multi-rank run)."""
        env, torch, model, B = self.env, self.env.torch, self.model, self.B
        pinned = self._pin()

        def hp(pb, k):
            return pb[k].ptr if k in pb else None
        if self.trainer is not None and env.args.exchange == "p2p":
            # peer-memory sharded step through its host-facing call: ps_model_p2p_submit copies this rank's slice from pinned host
            # memory and enqueues the step's graph, ps_model_collect returns the global loss of the oldest step in flight; the host
            # runs up to DEPTH steps ahead (every step's loss is still read back: Trainer.java:89 prints it per step)
            for i in range(n):
                pb = pinned[(start + i) % len(pinned)]
                model.p2p_submit_ptrs(hp(pb, "E"), hp(pb, "X"), hp(pb, "W"), hp(pb, "Y"), B)
                if mid is not None and i == n // 2:
                    mid()
                if i >= DEPTH - 1:
                    model.collect()
            for _ in range(min(n, DEPTH - 1) - 1):
                model.collect()
            return model.collect()
        if self.trainer is not None:      # NCCL sharded step: stage this rank's slice from pinned host memory, then the step, then its loss
            last = None
            for i in range(n):
                pb = pinned[(start + i) % len(pinned)]
                with torch.cuda.stream(env.stream):
                    d = {k: torch.from_numpy(pa.array).to(f"cuda:{env.local_rank}", non_blocking=True) for k, pa in pb.items()}
                self.trainer.step(d.get("E"), d["X"], d.get("W"), d["Y"])
                last = model.read_loss()                       # the step's result is read back every step
            return last
        for i in range(n):
            pb = pinned[(start + i) % len(pinned)]
            model.submit_ptrs(hp(pb, "E"), hp(pb, "X"), hp(pb, "W"), hp(pb, "Y"), B)
            if i >= DEPTH - 1:
                model.collect()
        for _ in range(min(n, DEPTH - 1) - 1):
            model.collect()
        return model.collect()

    def time_e2e(self, K, warmup, reps=None):
        env = self.env
        self.host_loop(len(self._pin()) + max(3, warmup), 0)        # every staging-buffer graph captured, untimed
        env.barrier()

        def one(start, sample=False):
            env.barrier()
            t0 = time.perf_counter()
            loss = self.host_loop(K, start, env.sampler.poll if (sample and env.sampler is not None) else None)
            env.ctx.synchronize()
            return (time.perf_counter() - t0) * 1e3, loss
        probe = env.max_over_ranks([one(warmup)[0]])[0]
        reps = reps or self._reps(K, probe)
        ms, loss = [], None
        for r in range(reps):
            t, loss = one(warmup + (r + 1) * K, sample=(r % 4 == 0))     # (one ~20 us NVML query inside every fourth repetition of ~4 ms)
            ms.append(t)
        ms = sorted(env.max_over_ranks(ms))
        h2d = sum(pa.nbytes for pa in self._pin()[0].values())
        return {"ms": ms[len(ms) // 2], "ms_min": ms[0], "ms_max": ms[-1], "reps": reps, "h2d": h2d, "loss": loss}

    def check(self):
        if self.trainer is not None:
            self.trainer.check()

    def unique_stats(self):
        F = self.cfg["F"]
        if not F:
            return 0.0, 0
        off = (np.arange(F, dtype=np.int64) << 44)[None, :]
        per = [len(np.unique(b["E"] + off)) for b in self.ring]
        allk = len(np.unique(np.concatenate([(b["E"] + off).ravel() for b in self.ring])))
        return float(np.mean(per)), allk

    def close(self):
        if self.pinned:
            for pb in self.pinned:
                for pa in pb.values():
                    pa.free()
            self.pinned = None
        self.dev_ring = None
        self.model.close()


def emb_rooflines(env, wl, hbm_peak, reps):
    """Device time of the embedding kernels replayed in a CUDA graph over the ring vs their algorithmic bytes."""
    cfg, B = wl.cfg, wl.B
    F, D = cfg["F"], cfg["D"]
    L = B * F
    uniq, ring_unique = wl.unique_stats()
    kt = wl.model.kernel_times([d["E"].data_ptr() for d in wl.dev_ring], B, reps=reps)
    alg = {"emb_lookup": L * (8 + 8 * D), "emb_scatter_update": L * (8 + 4 * D) + uniq * 24 * D}
    out = {}
    for k, bytes_ in alg.items():
        us = kt[k]
        out[k] = {"us": us, "alg_bytes": bytes_, "gbs": bytes_ / max(us, 1e-3) / 1e3, "frac": bytes_ / max(us, 1e-3) / 1e3 / hbm_peak}
    out["emb_resolve_only_us"] = kt["emb_resolve"]
    out["unique_keys_per_batch"] = uniq
    Dp = (D + 3) // 4 * 4
    ws = ring_unique * (12 * Dp + 16)
    out["ring_working_set_mb"] = ws / 1e6
    out["working_set_exceeds_l2"] = bool(ws > L2_BYTES)
    return out


def hbm_roofline(kernels, hbm_peak, peak_src, args):
    """The `roofline` object of the JSON line: the slower of the two embedding kernels (SURVEY 8d's algorithmic bytes / device time)."""
    dom = max(("emb_lookup", "emb_scatter_update"), key=lambda k: kernels[k]["us"])
    traffic = None                  # dram__bytes_read+write per launch from the committed ncu --set full capture of this round
    tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tp) and args.config == "cfg2":
        tk = json.load(open(tp)).get("kernels", {})
        # the scatter and the update each have a shared-memory slab variant (the default) and a register variant
        alts = {"emb_lookup": [["emb_lookup_kernel"]],
                "emb_scatter_update": [["emb_scatter_slab_kernel", "emb_scatter_kernel"], ["emb_update_slab_kernel", "emb_update_kernel"]]}[dom]
        picked = [next((p for p in names if p in tk), None) for names in alts]
        traffic = sum(tk[p]["dram_bytes_per_launch"] for p in picked) if all(picked) else None
    return {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": hbm_peak, "unit": "GB/s",
            "frac": kernels[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
            "how": "algorithmic bytes (SURVEY 8d) / mean device time of the kernel(s) replayed 64x in a CUDA graph over the batch ring "
                   "(CUDA events on the library's stream); emb_lookup = key resolution + gather in one kernel; emb_scatter_update = "
                   "the scatter launch + its programmatic dependent, the update launch"}


def large_batch_roofline(env, hbm_peak, name, seed):
    """The embedding kernels at a batch large enough to be bandwidth- rather than launch-latency-bound (cfg4 shapes: B=16384, D=64,
    10 M keys), for zipf AND uniform keys, over a ring of non-repeating batches whose rows exceed L2."""
    ps, torch = env.ps, env.torch
    out = {"peak": hbm_peak, "unit": "GB/s"}
    for dist_name in ("zipf", "uniform"):
        cfg = dict(CONFIGS[name])
        saved = env.args.dist
        env.args.dist = dist_name
        try:
            wl = Workload(env, name, cfg, ring=8, seed_off=77)
        finally:
            env.args.dist = saved
        for _ in range(2):                       # create the ring's keys, reach the steady state of the table
            for i in range(len(wl.dev_ring)):
                wl.dev_step(i)
        wl.model.read_loss()
        r = emb_rooflines(env, wl, hbm_peak, reps=32)
        r["workload"] = f"{name} shapes: B={cfg['B']} F={cfg['F']} D={cfg['D']} vocab={cfg['V']} keys={dist_name}, {r['unique_keys_per_batch']:.0f} unique keys/batch"
        out[dist_name] = r
        wl.close()
    return out


def tf32_ceiling(env):
    """cuBLAS TF32 GEMM 8192^3 (reference only — the denominator of the FcLayer roofline; BASELINE.md §2 asks for it)."""
    torch = env.torch
    torch.backends.cuda.matmul.allow_tf32 = True
    n = 8192
    a = torch.randn(n, n, device=f"cuda:{env.local_rank}")
    b = torch.randn(n, n, device=f"cuda:{env.local_rank}")
    for _ in range(3):
        torch.matmul(a, b)
    best = 1e9
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(10):
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def gemm_report(wl, tf32_peak, mma_per_product):
    cfg, B = wl.cfg, wl.B
    gt = wl.model.gemm_times(B, reps=64)
    dims = [cfg["F"] * cfg["D"] + cfg["Xn"]] + list(cfg["fc"])
    out = {}
    for l in range(len(cfg["fc"])):
        for n in ("forward", "dgrad", "wgrad"):
            us = gt[f"fc{l}.{n}"]
            fl = 2.0 * B * dims[l] * dims[l + 1]
            out[f"fc{l}.{n}"] = {"us": us, "tflops_fp32_equiv": fl / max(us, 1e-3) / 1e6, "mma_tflops": mma_per_product * fl / max(us, 1e-3) / 1e6,
                                 "frac_of_tf32_peak": (mma_per_product * fl / max(us, 1e-3) / 1e6) / tf32_peak if dims[l + 1] > 1 else None}
    return out


def parity_check(env, cfg, steps=2):
    """N-rank result == oracle(thread = 1, concatenated batch): PServer sync-mode semantics (PServer.java:164-214)."""
    ps, torch, dist = env.ps, env.torch, env.dist
    import oracle_lib as ol
    world, rank = env.world, env.rank
    Np = min(cfg["B"], 1024)
    pcfg = dict(cfg, B=Np)
    F, D, Xn, fc = cfg["F"], cfg["D"], cfg["Xn"], cfg["fc"]
    model = ps.Model(env.ctx, cfg["kind"], F, D, Xn, fc, emb_capacity=1 << 18, emb_updater=ps.UpdaterSpec.ftrl() if cfg["emb_opt"] == "ftrl" else None, max_batch=Np)
    trainer = None
    if env.sharded:
        from ps_b200.sharded import GpuOps, GraphedShardedTrainer, P2PShardedTrainer
        if env.args.exchange == "p2p":
            trainer = P2PShardedTrainer(ps, env.ctx, model, rank, world, Np, F, slack=3.0, device=env.local_rank)
        else:
            trainer = GraphedShardedTrainer(GpuOps(ps, env.ctx, model, env.local_rank), rank, world, Np, F, cfg["kind"] == "widedeep", slack=3.0)
    syn = Synth(F=F, Xn=Xn, V=min(cfg["V"], 200000), dist="zipf", seed=4242)
    batches = [syn.batch(world * Np) for _ in range(steps)]
    losses = []
    for b in batches:
        sl = slice(rank * Np, (rank + 1) * Np)
        if trainer is None:
            losses.append(model.train_step(b.get("E")[sl] if F else None, b["X"][sl], b.get("W")[sl] if cfg["kind"] == "widedeep" else None, b["Y"][sl]))
        else:
            d = {k: torch.from_numpy(np.ascontiguousarray(v[sl])).cuda(env.local_rank) for k, v in b.items()}
            trainer.step(d.get("E"), d["X"], d.get("W") if cfg["kind"] == "widedeep" else None, d["Y"])
            losses.append(model.read_loss())
    if trainer is not None:
        trainer.check()
    # sampled embedding keys of the last batch: each lives on exactly one rank
    keys = []
    if F:
        b = batches[-1]
        stride = max(1, (world * Np * F) // 400)
        flat = [(n, j) for n in range(world * Np) for j in range(F)][::stride][:400]
        keys = sorted({ol.key_string(0, j, int(b["E"][n, j])) for n, j in flat})
    mine = {}
    for k in keys:
        w = model.get(k)
        if w is not None:
            mine[k] = w
    dense = {f"fc{l}.weights": model.get(f"fc{l}.weights") for l in range(len(fc))}
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
    else:
        gathered = [mine]
    model.close()
    if rank != 0:
        return None
    o = oracle_model(pcfg, 20261017)
    lo = [o.train_step(b.get("E"), b["X"], b.get("W") if cfg["kind"] == "widedeep" else None, b["Y"]) for b in batches]
    loss_err = max(abs(a - b) / max(1e-12, abs(b)) for a, b in zip(losses, lo))
    tol_loss, tol_rows = (1e-4, 2e-3) if env.args.precision != "tf32" else (2e-3, 5e-2)
    rows, dup, rels = 0, 0, []
    seen = set()
    for g in gathered:
        for k, w in g.items():
            dup += k in seen
            seen.add(k)
            wo = o.get(k)
            rows += 1
            rels.append(float(np.abs(w - wo).max() / max(1e-6, float(np.abs(wo).max()))))
    rels = np.sort(np.asarray(rels)) if rels else np.zeros(1)
    # Adam's first steps move an element by ~alfa * sign(g) whatever |g| is: a gradient element within rounding of zero can
    # flip sign against the oracle's summation order and cost 2 * alfa on that element — tolerated on <= 1 % of the rows
    outside = int((rels > tol_rows).sum())
    max_rel = float(rels[-1])
    # dense weights: the same Adam sensitivity (an element whose gradient cancels to rounding noise moves by +-alfa per step in a
    # direction the summation order decides): at most 0.1 % of a matrix's elements may be outside the tolerance, none by more than
    # the 2 * alfa * steps an opposite-sign step sequence can produce
    dense_rel, dense_frac, dense_abs = 0.0, 0.0, 0.0
    for k, v in dense.items():
        ov = o.get(k)
        err = np.abs(v - ov)
        scale = max(1e-12, float(np.abs(ov).max()))
        dense_rel = max(dense_rel, float(np.quantile(err, 0.999)) / scale)
        dense_frac = max(dense_frac, float((err > tol_rows * scale).mean()))
        dense_abs = max(dense_abs, float(err.max()))
    ok = (loss_err <= tol_loss and outside <= max(1, rows // 100) and max_rel <= 2e-2 and dense_rel <= tol_rows and dense_frac <= 1e-3
          and dense_abs <= 2.2 * 0.005 * steps and rows >= min(200, len(keys)) and rows == len(keys) and dup == 0)
    return {"ok": bool(ok), "loss_err": loss_err, "rows_checked": rows, "rows_sampled": len(keys), "max_rel": max_rel, "rows_outside_tol": outside,
            "median_rel": float(rels[len(rels) // 2]), "dense_p999_rel": dense_rel, "dense_frac_outside_tol": dense_frac, "dense_max_abs": dense_abs,
            "keys_on_two_shards": dup, "steps": steps, "global_batch": world * Np, "tolerance": {"loss_rel": tol_loss, "rows_rel_to_max": tol_rows},
            "against": "CPU oracle, one Trainer step (thread = 1) per global batch on the concatenated batch"}


def extra_config(env, name, K, warmup):
    cfg = dict(CONFIGS[name])
    wl = Workload(env, name, cfg, ring=4, seed_off=300)
    wl.prepare()
    v = wl.time_value(K, max(3, warmup), reps=7)
    e = wl.time_e2e(K, max(3, warmup), reps=5)
    wl.check()
    total = cfg["B"] * env.world * K
    out = {"workload": workload_name(env.args, cfg, name), "global_batch": cfg["B"] * env.world, "value": total / (v["ms"] / 1e3), "unit": UNIT,
           "ms_per_step": v["ms"] / K, "e2e": total / (e["ms"] / 1e3), "reps": [v["reps"], e["reps"]], "loss": v["loss"],
           "table_rows_capacity_per_gpu": wl.cap, "table_gb_per_gpu": wl.cap * (16 + 12 * ((cfg["D"] + 3) // 4 * 4)) / 1e9}
    wl.close()
    return out


def time_ingest(env, cfg, batch, threads):
    """libsvm text → training: (a) the native host reader on `threads` host cores; (b) ps_model_submit_text — the raw text
    is copied to the GPU, parsed THERE and trained on in the same submission (DataSet.next + Trainer.train, DataSet.java:77-100)."""
    import tempfile
    ps = env.ps
    F, Xn, B = cfg["F"], cfg["Xn"], cfg["B"]
    if not F or cfg["kind"] != "widedeep":
        return None
    rows = min(B, 4096)
    lines = []
    for n in range(rows):
        cols = ["%d" % int(batch["Y"][n])] + ["%d:1" % int(batch["E"][n, j]) for j in range(F)] + ["%d:%.2f" % (33895 + x, batch["X"][n, x]) for x in range(Xn)]
        lines.append(" ".join(cols))
    text = ("\n".join(lines) + "\n").encode()
    out = {"bytes_per_line": len(text) / rows}
    with tempfile.NamedTemporaryFile("wb", suffix=".libsvm", delete=False) as f:
        f.write(text)
        path = f.name
    try:
        r = ps.LibsvmReader(path, F=F, Xn=Xn, batch=rows, threads=threads)
        bufs = dict(E=np.empty((rows, F), np.int64), W=np.empty((rows, F), np.int64), X=np.empty((rows, Xn), np.float32), Y=np.empty(rows, np.float32))
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 1.0:
            b = r.next(bufs)
            if b is None:
                r.reset()
                continue
            n += len(b["Y"])
        dt = time.perf_counter() - t0
        ok = bool(np.array_equal(bufs["E"][:rows], batch["E"][:rows]) and np.array_equal(bufs["X"][:rows], batch["X"][:rows]))
        r.close()
        out["host_reader"] = {"lines_per_s": n / dt, "threads": threads, "roundtrip_exact": ok}
    finally:
        os.unlink(path)
    # text → train on the GPU
    model = ps.Model(env.ctx, cfg["kind"], F, cfg["D"], Xn, cfg["fc"], emb_capacity=1 << 20, max_batch=rows)
    buf = ps.PinnedArray((len(text),), np.uint8)
    buf.array[:] = np.frombuffer(text, np.uint8)
    for i in range(4):
        model.submit_text(buf.ptr, len(text), rows)
        if i >= 1:
            model.collect()
    model.collect()
    info = model.step_info()
    steps = 200
    env.ctx.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        model.submit_text(buf.ptr, len(text), rows)
        if i >= 1:
            model.collect()
    model.collect()
    env.ctx.synchronize()
    dt = time.perf_counter() - t0
    out["text_to_train"] = {"lines_per_s": steps * rows / dt, "api": "ps_model_submit_text / ps_model_collect (2 steps in flight): H2D of the raw text, GPU parse, train step",
                            "h2d_bytes_per_step": len(text), "bad_lines": info["bad_lines"], "skipped": info["skipped"], "batch": rows}
    buf.free()
    model.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--dist", default="zipf")
    # FcLayer arithmetic: tf32x3 = tcgen05 tensor cores with error-compensated operand split (fp32-grade results, the library
    # default: the reference computes in fp32); tf32 = plain TF32 tensor cores; fp32 = FFMA exact mode
    ap.add_argument("--precision", default=os.environ.get("PS_FC_PRECISION", "tf32x3"))
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--ring", type=int, default=64)
    ap.add_argument("--slack", type=float, default=1.25, help="per-owner bucket capacity = lookups / ranks * slack (de-duplicated keys need far less; uniform keys ~1.05)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--force-sharded", action="store_true", help="use the sharded step even with one rank (profiling)")
    ap.add_argument("--roofline-at-n", action="store_true", help="N > 1: rank 0 also times the two embedding kernels on a standalone table (the N = 1 roofline)")
    ap.add_argument("--large", default="cfg4", help="shapes for the large-batch embedding roofline ('' = skip)")
    ap.add_argument("--extra", default="cfg3,cfg4", help="further BASELINE configs measured at this N ('' = skip)")
    ap.add_argument("--no-kernel-times", action="store_true", help="skip the per-kernel replays and side measurements (ncu runs: only the step's own launches)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--reps", type=int, default=0, help="repetitions of the K-step loop (0 = at least 25 and at least 1 s)")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    rank = int(os.environ.get("RANK", "0"))

    if args.impl == "reference":
        import __graft_entry__ as g
        if rank == 0:
            g.build()
            run_reference(args, cfg)
        return

    # The cyclic collector never runs by itself in this process: started from the clock-sampler thread (which allocates ctypes objects a few
    # hundred times a second) it would run finalizers while the main thread is inside a native call — two 8-GPU runs ended in a segmentation
    # fault inside exactly such a collection (faulthandler: "Garbage-collecting" under nvmlDeviceGetClockInfo, the main thread in
    # ps_model_p2p_submit).  With one rank it is run explicitly, in the main thread, between the measurements (which also keeps GC pauses out
    # of the timed loops); a multi-rank run, which leaves through os._exit, does not run it at all — everything large is released explicitly.
    import faulthandler
    import gc
    faulthandler.enable()            # a crash in any rank leaves its Python stacks in the log
    gc.disable()
    env = Env(args)
    collect = gc.collect if env.world == 1 else (lambda: 0)
    world, B, F, D = env.world, cfg["B"], cfg["F"], cfg["D"]
    side = not args.no_kernel_times
    hbm_peak, peak_src = peaks()

    # ---- the headline workload ----
    wl = Workload(env, args.config, cfg, ring=args.ring)
    wl.prepare()
    collect()
    sampler = ClockSampler(env.local_rank, threaded=(env.world == 1))
    env.sampler = sampler
    sampler.begin()
    v = wl.time_value(args.steps, args.warmup, reps=args.reps or None)
    e = wl.time_e2e(args.steps, args.warmup, reps=args.reps or None)
    clocks = sampler.summary()
    env.sampler = None
    collect()
    wl.check()

    # ---- per-kernel device times and rooflines (one GPU, local step) ----
    kernels, roofline, roofline_tensor, large, tf32_peak, phase_us, cfg5 = {}, None, None, None, None, {}, None
    if side and (wl.trainer is None or args.exchange == "p2p"):
        # events between the kernels of the MAIN stream (the step's critical path); at N > 1 every rank steps, rank 0's times are reported
        acc = {}
        wl.model.profile(True)
        for i in range(20):
            wl.dev_step(i)
            wl.model.read_loss()
            for k, t in wl.model.phase_times().items():
                acc.setdefault(k, []).append(t)
        wl.model.profile(False)
        phase_us = {k: 1e3 * float(np.median(t)) for k, t in acc.items()}
    if side and world == 1 and wl.trainer is None:
        tf32_peak = tf32_ceiling(env)
        mma_per = {"fp32": 0, "tf32": 1, "tf32x3": 3}[args.precision]
        if F:
            kernels = emb_rooflines(env, wl, hbm_peak, reps=64)
            roofline = hbm_roofline(kernels, hbm_peak, peak_src, args)
        gk = gemm_report(wl, tf32_peak, max(mma_per, 1))
        kernels.update(gk)
        big = max((k for k in gk if gk[k]["frac_of_tf32_peak"] is not None), key=lambda k: gk[k]["us"])
        roofline_tensor = {"bound": "tensor", "kernel": f"gemm_tf32_kernel ({big})", "achieved": gk[big]["mma_tflops"], "peak": tf32_peak, "unit": "TFLOP/s",
                           "frac": gk[big]["frac_of_tf32_peak"], "fp32_equivalent_tflops": gk[big]["tflops_fp32_equiv"],
                           "peak_source": "cuBLAS TF32 8192^3 (torch.matmul, allow_tf32), best of 10, measured in this run",
                           "how": f"{max(mma_per, 1)} TF32 MMAs per product ({args.precision}); device time of the GEMM replayed 64x in a CUDA graph"}
        if args.large:
            large = large_batch_roofline(env, hbm_peak, args.large, 20261017 + 4)
        c5 = dict(CONFIGS["cfg5"])
        w5 = Workload(env, "cfg5", c5, ring=4, seed_off=500)
        w5.prepare()
        v5 = w5.time_value(min(args.steps, 50), 3, reps=7)
        cfg5 = {"workload": workload_name(args, c5, "cfg5"), "value": c5["B"] * min(args.steps, 50) / (v5["ms"] / 1e3), "ms_per_step": v5["ms"] / min(args.steps, 50),
                "gemms": gemm_report(w5, tf32_peak, max(mma_per, 1))}
        w5.close()

    ingest = None
    if side and world == 1 and wl.trainer is None and rank == 0:
        try:
            ingest = time_ingest(env, cfg, wl.ring[0], min(16, os.cpu_count() or 1))
        except Exception as ex:        # host-side plumbing: never fail the bench line over it
            ingest = {"error": str(ex)}

    uniq, ring_unique = wl.unique_stats()
    nvlink = getattr(wl, "nvlink", None)
    if nvlink is None and env.world > 1:
        nvlink = {"unavailable": "; ".join(NVLINK_WHY) or "counters did not move"}
    if nvlink is not None and F:
        Dp_ = (D + 3) // 4 * 4
        remote = uniq * (world - 1) / world          # unique keys of a batch owned by another rank
        glen = sum((a + 1) * b for a, b in zip([F * D + cfg["Xn"]] + list(cfg["fc"])[:-1], cfg["fc"])) + 3
        nvlink["expected_payload_bytes_per_step"] = (remote * (8 + 2 * 4 * Dp_ + 4)            # keys out, rows back, gradient sums + counts out
                                                     + (world - 1) * (4 * glen + (8 * B * F if cfg["kind"] == "widedeep" else 0)))   # dense sums, wide ids to every replica
        nvlink["how"] = "NVML NVLINK_THROUGHPUT_DATA_TX/RX of rank 0's GPU around the timed device-resident steps; expected = de-duplicated keys, rows and gradient sums to/from other owners + dense gradient sums and wide ids to every replica"
    if side and world > 1 and rank == 0 and F and args.roofline_at_n:
        # (opt-in: --roofline-at-n; verified at N = 2 only)  the same two embedding kernels on rank 0's GPU, on a standalone single-GPU table holding the whole vocabulary, over rank 0's ring
        # (the sharded step runs these kernels' owner / requester forms; this is the number the N = 1 line reports)
        class _Local:
            pass
        lw = _Local()
        lw.cfg, lw.B, lw.dev_ring, lw.unique_stats = cfg, B, wl.dev_ring, wl.unique_stats
        upd = env.ps.UpdaterSpec.ftrl() if cfg["emb_opt"] == "ftrl" else None
        lw.model = None
        try:
            lw.model = env.ps.Model(env.ctx, cfg["kind"], F, D, cfg["Xn"], cfg["fc"], emb_capacity=int(min(2 ** 31 - 1, 2 * cfg["V"] + (1 << 16))),
                                    emb_updater=upd, max_batch=B)
            kernels = emb_rooflines(env, lw, hbm_peak, reps=64)
            roofline = hbm_roofline(kernels, hbm_peak, peak_src, args)
            roofline["how"] += "; N > 1: measured on rank 0 with a standalone single-GPU table"
        except Exception as ex:                       # a side measurement must not cost the line
            kernels, roofline = {"error": str(ex)}, None
        finally:
            if lw.model is not None:
                lw.model.close()

    loss_value = v["loss"]
    wl_cap = wl.cap
    wl_ring = len(wl.ring)
    ring0 = wl.ring[:4]
    wl.close()

    # ---- parity at this N, then the other BASELINE configs at this N ----
    parity = None
    collect()
    if not args.no_parity:
        parity = parity_check(env, cfg)
        collect()
    extras = {}
    if side and args.extra:
        for name in [x for x in args.extra.split(",") if x]:
            try:
                extras[name] = extra_config(env, name, min(args.steps, 20), 3)
                collect()
            except Exception as ex:
                extras[name] = {"error": str(ex)}
                if world > 1:
                    raise                 # ranks must not diverge inside a sharded step

    if rank == 0:
        cpu = None
        if world == 1 and side:
            sps, n, gemm, dt = time_oracle_replicas(cfg, ring0, args.cpu_budget, 1)
            cpu = {"value": sps, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"{n} Trainer steps of batch {B} in {dt:.1f}s, thread=1 (CTR.java:72); C++ restatement of the reference's standalone "
                             f"Java path (no JVM in this image); sgemm={gemm}"}
        total = B * world * args.steps
        Dp = (D + 3) // 4 * 4
        line = {
            "metric": METRIC, "value": total / (v["ms"] / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": v["ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "tf32": "tf32", "tf32x3": "f32 via 3xTF32 tcgen05"}[args.precision], "data": "synthetic",
            "config": {"workload": workload_name(args, cfg), "global_batch": B * world},
            "config_detail": {
                "parallelism": (f"key-hash sharded embedding table over {world} GPUs, exchange={args.exchange} "
                                f"({'NVLink peer-memory stores with in-kernel flags, no collective calls' if args.exchange == 'p2p' else 'NCCL all-to-all'}) "
                                "+ data-parallel dense") if world > 1 else "single",
                "l2": ("a ring of %d distinct batches per GPU touching %.0f MB of table rows + slot records (%s the 126 MB L2); table %.0f MB; no flush" % (
                    wl_ring, ring_unique * (12 * Dp + 16) / 1e6, "exceeds" if ring_unique * (12 * Dp + 16) > L2_BYTES else "FITS IN", wl_cap * (16 + 12 * Dp) / 1e6)),
                "timing": "every ring batch stepped once untimed (graph capture, key inserts), then --warmup steps, then the K-step loop repeated; median repetition reported",
            },
            "timing": {"value_reps": v["reps"], "value_ms_min_med_max": [v["ms_min"], v["ms"], v["ms_max"]],
                       "e2e_reps": e["reps"], "e2e_ms_min_med_max": [e["ms_min"], e["ms"], e["ms_max"]],
                       "timed_region_s": (v["ms"] * v["reps"] + e["ms"] * e["reps"]) / 1e3},
            "e2e": {"value": total / (e["ms"] / 1e3), "unit": UNIT, "h2d_bytes_per_step": e["h2d"], "d2h_bytes_per_step": 32,
                    "api": f"ps_model_submit/ps_model_collect (up to {DEPTH} steps in flight, every step's loss read back)" if not env.sharded else
                           f"ps_model_p2p_submit/ps_model_collect (up to {DEPTH} steps in flight, every step's loss read back)" if args.exchange == "p2p" else
                           "pinned host batch -> device (async copy on the step's stream) -> sharded step -> ps_model_read_loss, every step"},
            "gpu_launches": int(round(v["launches"])), "clocks": clocks, "roofline": roofline, "roofline_tensor": roofline_tensor,
            "roofline_large_batch": large, "tf32_peak_tflops": tf32_peak, "kernels_us": phase_us, "hbm_kernels": kernels, "cfg5": cfg5,
            "parity": parity, "extra_configs": extras, "nvlink": nvlink,
            "cpu_baseline": cpu, "ingest": ingest, "loss": loss_value, "loss_e2e": e["loss"], "unique_keys_per_batch": uniq,
        }
        print(json.dumps(line), flush=True)
    sys.stdout.flush()
    if env.sharded:
        # captured graphs + communicator teardown order is fragile: everything is measured and printed, leave hard
        env.barrier()
        os._exit(0)
    env.ctx.close()


if __name__ == "__main__":
    main()
