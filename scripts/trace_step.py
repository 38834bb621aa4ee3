#!/usr/bin/env python
"""Kernel timeline of the training step (CUPTI through torch.profiler; there is no nsys in the image).

  python scripts/trace_step.py [--config cfg2] [--precision tf32x3] [--steps 6] [--out gpurun_out/trace.json]

Writes one JSON list of {name, ts_us, dur_us, stream} for every kernel of the profiled steps, in start order, and
prints a per-step table: start offset of each kernel relative to the step's first kernel, duration, stream.
A number measured under the profiler is never a bench value; this is for finding gaps in the step graph.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--precision", default="tf32x3")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "trace.json"))
    args = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile

    import __graft_entry__ as g
    g.build()
    from ps_b200 import binding as ps
    from ps_b200.synth import CONFIGS, Synth
    cfg = CONFIGS[args.config]
    B, F, D, Xn = cfg["B"], cfg["F"], cfg["D"], cfg["Xn"]
    ctx = ps.Context(0, seed=20261017)
    ctx.set_fc_precision({"fp32": ps.PS_FC_FP32, "tf32": ps.PS_FC_TF32, "tf32x3": ps.PS_FC_TF32X3}[args.precision])
    model = ps.Model(ctx, cfg["kind"], F, D, Xn, cfg["fc"], emb_capacity=2 * cfg["V"] + (1 << 16) if cfg["V"] else 1024, max_batch=B)
    syn = Synth(F=F, Xn=Xn, V=cfg["V"], seed=20261017 + 2, n_classes=10 if cfg["kind"] == "fcnn" else 0)
    ring = [{k: torch.from_numpy(np.ascontiguousarray(v)).cuda(0) for k, v in syn.batch(B).items()} for _ in range(8)]
    torch.cuda.synchronize()

    def p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def step(i):
        d = ring[i % len(ring)]
        model.train_step_dev(p(d.get("E")), p(d["X"]), p(d.get("W")), p(d["Y"]), B)
    for i in range(24):
        step(i)
    model.read_loss()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(args.steps):
            step(24 + i)
        model.read_loss()
    tmp = args.out + ".chrome.json"
    prof.export_chrome_trace(tmp)
    ev = json.load(open(tmp))["traceEvents"]
    os.remove(tmp)
    ks = sorted(({"name": e["name"], "ts_us": e["ts"], "dur_us": e["dur"], "stream": e.get("args", {}).get("stream")}
                 for e in ev if e.get("cat") == "kernel"), key=lambda k: k["ts_us"])
    json.dump(ks, open(args.out, "w"))
    # split into steps at the probe kernel (first kernel of a step with an embedding layer) or the first kernel name seen
    first = next((k["name"] for k in ks if "probe" in k["name"]), ks[0]["name"] if ks else "")
    starts = [i for i, k in enumerate(ks) if k["name"] == first]
    for si in range(max(0, len(starts) - 2), len(starts)):
        a, b = starts[si], starts[si + 1] if si + 1 < len(starts) else len(ks)
        t0 = ks[a]["ts_us"]
        print(f"--- step {si}: {b - a} kernels, span {ks[b - 1]['ts_us'] + ks[b - 1]['dur_us'] - t0:.1f} us"
              + (f", next step starts at +{ks[b]['ts_us'] - t0:.1f} us" if b < len(ks) else ""))
        for k in ks[a:b]:
            print(f"  +{k['ts_us'] - t0:8.1f}  {k['dur_us']:7.1f} us  s{k['stream']}  {k['name'][:70]}")
    os._exit(0)


if __name__ == "__main__":
    main()
