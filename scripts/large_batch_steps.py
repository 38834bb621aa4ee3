"""Runs a few training steps at the large-batch embedding shapes (cfg4: B=16384, D=64, 10 M keys) so that ncu can capture the
embedding kernels where they are bandwidth- rather than latency-bound:
  ncu --set full --clock-control none --import-source on -k regex:emb_ -s 64 -c 8 -o gpurun_out/prof_large python scripts/large_batch_steps.py
(16 warm-up steps x 4 embedding launches are skipped by -s 64)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402
from ps_b200.synth import CONFIGS, Synth  # noqa: E402

g.build()
from ps_b200 import binding as ps  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sharded = len(sys.argv) > 3 and sys.argv[3] == "sharded"      # the peer-memory sharded step on ONE rank (every exchange kernel runs, no link)
cfg = dict(CONFIGS[name])
B, F, D, Xn, V = cfg["B"], cfg["F"], cfg["D"], cfg["Xn"], cfg["V"]
ctx = ps.Context(0, seed=20261017)
ctx.set_fc_precision(ps.PS_FC_TF32X3)
model = ps.Model(ctx, cfg["kind"], F, D, Xn, cfg["fc"], emb_capacity=2 * V + (1 << 16), max_batch=B)
syn = Synth(F=F, Xn=Xn, V=V, dist="zipf", seed=20261021)
ring = [{k: torch.from_numpy(np.ascontiguousarray(v)).cuda(0) for k, v in syn.batch(B).items()} for _ in range(8)]
torch.cuda.synchronize()


def p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


trainer = None
if sharded:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29544")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    from ps_b200.sharded import P2PShardedTrainer
    trainer = P2PShardedTrainer(ps, ctx, model, 0, 1, B, F, slack=2.0, device=0)
for i in range(16 + steps):
    d = ring[i % 8]
    if trainer is not None:
        trainer.step(d.get("E"), d["X"], d.get("W"), d["Y"])
    else:
        model.train_step_dev(p(d.get("E")), p(d["X"]), p(d.get("W")), p(d["Y"]), B)
print("loss", model.read_loss())
if trainer is not None:
    os._exit(0)
model.close()
ctx.close()
