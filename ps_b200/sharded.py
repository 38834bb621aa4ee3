"""Key-hash sharded training step across the GPUs of one box — the role of the reference's
net.PSRouterClient (PSRouterClient.java:60-151: bucket keys by router.shard(key), one batched
getList / updateList per shard in parallel, merge) with net.PServer's sync-mode semantics
(PServer.java:164-283: sum the pushes of all workers, update once), one process per GPU.

The R ranks together perform ONE train.Trainer step (thread = 1) on the concatenated batch:
an R-GPU step equals the 1-GPU step on the same global batch up to fp32 reassociation.

  rank-local kernels (C ABI, ps_*_shard_*)            collectives (torch.distributed / NCCL, NVLink)
  ------------------------------------------------   -----------------------------------------------
  route: owner = hash(key) mod R, bucket by owner      all_to_all(counts), all_to_all(keys)
  owner: find-or-insert + gather (+ReLU)               all_to_all(rows)            [getList]
  requester: unpack rows → concat buffer
  dense: wide + FcLayer fwd/bwd, flat grad buffer      all_gather(wide ids), all_reduce(grad buffer)
  requester: per-lookup row gradients (ReLU mask)      all_to_all(row gradients)   [push]
  owner: fused scatter-add + g_eff + Adam/Ftrl         (barrier = stream order)

`ops` supplies the rank-local pieces: GpuOps (the CUDA library) in production; the gloo CPU tests
drive the same orchestration with a numpy stand-in to check the routing / split arithmetic.
"""
import contextlib
import ctypes as C

import torch
import torch.distributed as dist


class ShardedTrainer:
    def __init__(self, ops, rank, world, group=None):
        self.ops, self.rank, self.world, self.group = ops, rank, world, group

    def _a2a(self, out, inp, out_splits=None, in_splits=None):
        dist.all_to_all_single(out, inp, out_splits, in_splits, group=self.group)
        return out

    def step(self, E, X, W, Y):
        """E, W: [N, F] int64, X: [N, Xn] f32, Y: [N] f32 — this rank's slice of the global batch, on
        the ops' device.  Returns the global loss (python float)."""
        with self.ops.scope():
            return self._step(E, X, W, Y)

    def _step(self, E, X, W, Y):
        ops, R = self.ops, self.world
        N = int(Y.shape[0])
        has_emb = E is not None
        if has_emb:
            send_keys, send_pos, counts = ops.route(E, R)
            send_counts = counts.to(torch.int64)
            recv_counts = torch.empty_like(send_counts)
            self._a2a(recv_counts, send_counts)
            sc, rc = send_counts.tolist(), recv_counts.tolist()          # the one host sync of the step
            recv_keys = ops.empty(sum(rc), torch.int64)
            self._a2a(recv_keys, send_keys, rc, sc)                     # PSRouterClient.getList: keys out
            rows_out = ops.lookup(recv_keys)                            # PServer.getList on the owner
            rows_back = ops.empty((sum(sc), rows_out.shape[1]), torch.float32)
            self._a2a(rows_back, rows_out, sc, rc)                      # rows back
            ops.unpack(rows_back, send_pos, N)
        W_all = None
        if W is not None and ops.has_wide:
            W_all = ops.empty((R * N,) + tuple(W.shape[1:]), torch.int64)
            dist.all_gather_into_tensor(W_all, W.contiguous(), group=self.group)
        ops.dense_step(X, W, W_all, Y, N)
        dist.all_reduce(ops.grad_buffer(), group=self.group)           # dense gradient sums + loss + gbar
        if has_emb:
            grads_send = ops.pack_grads(send_pos, N)
            grads_recv = ops.empty((sum(rc), grads_send.shape[1]), torch.float32)
            self._a2a(grads_recv, grads_send, rc, sc)                   # KVStore.update → client.push per key
        ops.finish(N * R, R)                                            # PServer.psUpdate for dense + wide keys
        if has_emb:
            ops.apply(grads_recv)                                       # ... and for the embedding rows this rank owns
        return ops.loss()


class GraphedShardedTrainer:
    """The sharded step with FIXED per-owner bucket capacity, captured once into a CUDA graph per rank
    (local kernels + the NCCL collectives) and replayed: no host synchronisation inside a step, one graph
    launch per step.  Buckets are padded with the EMPTY key, which owners skip.  If a bucket ever
    overflows (`overflow` flag, checked by check()), use ShardedTrainer (exact sizes) for that workload
    or raise `slack`."""

    def __init__(self, ops, rank, world, N, F, has_wide, group=None, slack=2.0):
        self.ops, self.rank, self.world, self.group, self.N, self.F, self.has_wide = ops, rank, world, group, N, F, has_wide
        L = N * F
        self.cap = int(-(-int(L / world * slack + 64) // 32) * 32) if world > 1 else L
        self.graph, self.static = None, None

    def _body(self, E, X, W, Y):
        ops, R, N, cap = self.ops, self.world, self.N, self.cap
        send_keys, send_pos = ops.route_padded(E, R, cap)
        recv_keys = ops.empty(R * cap, torch.int64)
        dist.all_to_all_single(recv_keys, send_keys, group=self.group)
        rows_out = ops.lookup(recv_keys)
        rows_back = ops.empty((R * cap, rows_out.shape[1]), torch.float32)
        dist.all_to_all_single(rows_back, rows_out, group=self.group)
        ops.unpack(rows_back, send_pos, N)
        W_all = None
        if W is not None and self.has_wide:
            W_all = ops.empty((R * N,) + tuple(W.shape[1:]), torch.int64)
            dist.all_gather_into_tensor(W_all, W, group=self.group)
        ops.dense_step(X, W, W_all, Y, N)
        dist.all_reduce(ops.grad_buffer(), group=self.group)
        grads_send = ops.pack_grads_padded(send_pos, N, R * cap)
        grads_recv = ops.empty((R * cap, grads_send.shape[1]), torch.float32)
        dist.all_to_all_single(grads_recv, grads_send, group=self.group)
        ops.finish(N * R, R)
        ops.apply(grads_recv)

    def step(self, E, X, W, Y, warmup_eager=2):
        """Enqueues one step (asynchronous).  The first calls run eagerly (allocations, NCCL warm-up), then
        the step is captured; afterwards each call copies the inputs into the graph's static buffers and
        replays.  Returns nothing: read the loss with ops.loss() when needed."""
        ops = self.ops
        with ops.scope():
            if self.static is None:
                self.static = {"E": E.clone(), "X": X.clone(), "W": None if W is None else W.clone(), "Y": Y.clone()}
                self.calls = 0
            st = self.static
            st["E"].copy_(E); st["X"].copy_(X); st["Y"].copy_(Y)
            if W is not None:
                st["W"].copy_(W)
            if self.graph is None and self.calls < warmup_eager:
                l0 = ops.ctx.launch_count()
                self._body(st["E"], st["X"], st["W"], st["Y"])
                self.launches_per_step = ops.ctx.launch_count() - l0      # this library's kernels per step (NCCL's not counted)
                self.calls += 1
                return
            if self.graph is None:
                ops.synchronize()
                dist.barrier(group=self.group)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=ops.stream, capture_error_mode="thread_local"):
                    self._body(st["E"], st["X"], st["W"], st["Y"])
                self.graph = g
            self.graph.replay()

    def check(self):
        if self.ops.overflowed():
            raise RuntimeError("sharded exchange bucket overflow: raise slack or use ShardedTrainer")


class P2PShardedTrainer:
    """The sharded step over NVLink peer memory (ps_b200/csrc/p2p.cu): torch.distributed is used ONCE, to
    exchange the CUDA-IPC handles of the mailbox slabs; afterwards a step is a single C call that
    replays one CUDA graph per rank — no collective library, no host synchronisation on the data path."""

    def __init__(self, ps, ctx, model, rank, world, N, F, group=None, slack=2.0, device=None):
        self.ps, self.ctx, self.model, self.rank, self.world, self.group = ps, ctx, model, rank, world, group
        self.lib = ps.lib()
        L = N * max(F, 1)
        self.cap = int(-(-int(L / world * slack + 64) // 32) * 32) if world > 1 else L
        dev = torch.device("cuda", rank if device is None else device)
        handle = (C.c_ubyte * 64)()
        ps.check(self.lib.ps_model_p2p_init(model.h, world, rank, self.cap, handle))
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
        allh = torch.empty(64 * world, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, mine, group=group)
        buf = (C.c_ubyte * (64 * world))(*allh.cpu().tolist())
        ps.check(self.lib.ps_model_p2p_connect(model.h, buf))
        dist.barrier(group=group)                 # every slab is mapped everywhere before the first store

    def step(self, E, X, W, Y):
        """Enqueues one step (asynchronous) on this rank's slice of the global batch (device tensors)."""
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        self.ps.check(self.lib.ps_model_p2p_step_dev(self.model.h, p(E), p(X), p(W), p(Y), int(Y.shape[0])))

    def loss(self):
        return self.model.read_loss()

    def check(self):
        v = C.c_int()
        self.ps.check(self.lib.ps_model_p2p_overflowed(self.model.h, C.byref(v)))
        if v.value:
            raise RuntimeError("p2p exchange bucket overflow: raise slack")


class _DevArray:
    def __init__(self, ptr, n, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class GpuOps:
    """Rank-local pieces on the CUDA library; every tensor lives on the library's device and every
    call is ordered on the library's stream, which is made torch's current stream so the NCCL
    collectives between the calls are ordered with them."""

    def __init__(self, ps, ctx, model, device):
        self.ps, self.ctx, self.model, self.device = ps, ctx, model, torch.device("cuda", device)
        self.lib = ps.lib()
        self.stream = torch.cuda.ExternalStream(ctx.stream(), device=self.device)
        self.has_wide = model.kind in ("widedeep", ps.PS_MODEL_WIDEDEEP)
        dp = C.c_int()
        if model.F:
            ps.check(self.lib.ps_model_shard_row_stride(model.h, C.byref(dp)))
        self.Dp = dp.value
        buf, cnt = C.c_void_p(), C.c_int64()
        ps.check(self.lib.ps_model_shard_grad_buffer(model.h, C.byref(buf), C.byref(cnt)))
        self._gbuf = torch.as_tensor(_DevArray(buf.value, cnt.value), device=self.device)

    def scope(self):
        return torch.cuda.stream(self.stream)

    @staticmethod
    def _ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else None

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def route(self, E, R):
        N, F = E.shape
        L = N * F
        send_keys, send_pos = self.empty(L, torch.int64), self.empty(L, torch.int32)
        counts, cursor = self.empty(R, torch.int32), self.empty(R, torch.int32)
        self.ps.check(self.lib.ps_shard_route_dev(self.ctx.h, self._ptr(E), N, F, R, self._ptr(send_keys), self._ptr(send_pos),
                                                  self._ptr(counts), self._ptr(cursor)))
        return send_keys, send_pos, counts

    def route_padded(self, E, R, cap):
        N, F = E.shape
        send_keys, send_pos = self.empty(R * cap, torch.int64), self.empty(N * F, torch.int32)
        if not hasattr(self, "_cursor"):
            self._cursor, self._overflow = self.empty(64, torch.int32), torch.zeros(1, dtype=torch.int32, device=self.device)
        self.ps.check(self.lib.ps_shard_route_padded_dev(self.ctx.h, self._ptr(E), N, F, R, cap, self._ptr(send_keys), self._ptr(send_pos),
                                                         self._ptr(self._cursor), self._ptr(self._overflow)))
        return send_keys, send_pos

    def overflowed(self):
        return hasattr(self, "_overflow") and bool(self._overflow.item())

    def synchronize(self):
        self.ctx.synchronize()
        torch.cuda.synchronize(self.device)

    def pack_grads_padded(self, send_pos, N, rows):
        g = self.empty((rows, self.Dp), torch.float32)
        self.ps.check(self.lib.ps_model_shard_pack_grads_dev(self.model.h, self._ptr(send_pos), N, self._ptr(g)))
        return g

    def lookup(self, keys):
        n = keys.numel()
        rows = self.empty((n, self.Dp), torch.float32)
        self.ps.check(self.lib.ps_model_shard_lookup_dev(self.model.h, self._ptr(keys), n, self._ptr(rows)))
        return rows

    def unpack(self, rows, send_pos, N):
        self.ps.check(self.lib.ps_model_shard_unpack_dev(self.model.h, self._ptr(rows), self._ptr(send_pos), N))

    def dense_step(self, X, W, W_all, Y, N):
        n_all = 0 if W_all is None else W_all.numel()
        self.ps.check(self.lib.ps_model_shard_dense_step_dev(self.model.h, self._ptr(X), self._ptr(W), self._ptr(W_all), n_all, self._ptr(Y), N))

    def grad_buffer(self):
        return self._gbuf

    def pack_grads(self, send_pos, N):
        g = self.empty((send_pos.numel(), self.Dp), torch.float32)
        self.ps.check(self.lib.ps_model_shard_pack_grads_dev(self.model.h, self._ptr(send_pos), N, self._ptr(g)))
        return g

    def finish(self, N_global, R):
        self.ps.check(self.lib.ps_model_shard_finish_dev(self.model.h, N_global, R))

    def apply(self, grads):
        self.ps.check(self.lib.ps_model_shard_apply_dev(self.model.h, self._ptr(grads), grads.shape[0]))

    def loss(self):
        return self.model.read_loss()
