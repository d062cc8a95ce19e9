"""Multi-GPU LightGCN: users range-sharded, items replicated, one all-reduce of the [I, d] item block per
propagation layer (SURVEY.md 8e "bipartite-aware 1.5-D row sharding").  One process per GPU, NCCL through
``torch.distributed`` (the reference has no distributed code at all: this layer is new).

Per rank g (owning users [u_g, u_{g+1}), split at nnz-balanced points):
    local table   [U_g + I, d]   rows 0..U_g-1 = owned user embeddings, rows U_g.. = ALL item embeddings (replicated)
    local graph   the symmetric block [[0, R_g], [R_g^T, 0]] with the GLOBAL symmetric normalisation
                  (user degrees are local, item degrees are all-reduced once at build time)
    two row views of that CSR (free: rowptr slices)
        G_users : rows 0..U_g-1      users <- items   complete locally (items are replicated)
        G_items : rows U_g..U_g+I-1  items <- users   PARTIAL sums over the owned users only
A layer is   Y_items(partial) = G_items X   ->  all-reduce(Y_items) on the comm stream
          || Y_users = G_users X with the fused accumulate epilogue on the compute stream   (overlap)
          -> item rows: acc += all-reduced Y_items (small [I, d] pass).
User embeddings never move.  The backward is the same pattern on the (symmetric) block with the
residual r = dE_f/(K+1); BPR triples are handled by the rank that owns the user, item-row gradient
contributions are partial sums folded into the all-reduces.

The collective / kernel calls go through a small ``ops`` object so that the partition + exchange logic
can be exercised on the CPU with gloo and an oracle-backed ops object in tests/ (the product ops object
is CUDA-only and fails loudly otherwise).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, ptr, stream
from .bpr import LgbBprArgs, _launch as _bpr_launch, _ws as _bpr_ws
from .csr import DEFAULT_CHUNK, DeviceCSR


# --------------------------------------------------------------------------------------------
# partitioning (pure index arithmetic; shared by the CUDA path and the CPU tests)
# --------------------------------------------------------------------------------------------
def balanced_user_bounds(user_degree: torch.Tensor, world: int) -> List[int]:
    """Split points u_0=0 <= u_1 <= ... <= u_G=U so every range carries ~nnz/G interactions."""
    U = user_degree.numel()
    csum = torch.cumsum(user_degree.to(torch.int64), 0)
    total = int(csum[-1]) if U else 0
    bounds = [0]
    for g in range(1, world):
        target = (total * g) // world
        b = int(torch.searchsorted(csum, torch.tensor(target, device=csum.device), right=False)) + 1 if U else 0
        bounds.append(min(max(b, bounds[-1]), U))
    bounds.append(U)
    return bounds


def local_block(users: torch.Tensor, items: torch.Tensor, lo: int, hi: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """COO of the local symmetric block in local numbering (users 0..U_g-1, items U_g..U_g+I-1)."""
    m = (users >= lo) & (users < hi)
    lu = users[m] - lo
    li = items[m] + (hi - lo)
    return torch.cat([lu, li]), torch.cat([li, lu]), lu, items[m]


# --------------------------------------------------------------------------------------------
# CUDA ops object
# --------------------------------------------------------------------------------------------
class CudaOps:
    """liblaplace_b200 kernels + NCCL collectives.  ``comm`` stream carries the all-reduces."""

    def __init__(self, device: torch.device, group=None):
        if device.type != "cuda":
            raise RuntimeError("ShardedLightGCN needs CUDA devices (no CPU fallback); tests inject their own ops object")
        self.device, self.group = device, group
        self.comm = torch.cuda.Stream(device=device)
        self._bpr_ws = None
        self.exchange_events = None     # set to a list to record (start, end, bytes) CUDA events of every item-block exchange

    # graph ----------------------------------------------------------------------------------
    def build_graph(self, row, col, n, dinv):
        g = DeviceCSR.from_coo(row, col, n, n, chunk=0)
        val = torch.empty(g.nnz, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(_lib.load().lgb_gcn_values(ptr(g.rowptr), ptr(g.colidx), n, g.nnz, ptr(dinv), ptr(val), stream()), "gcn_values")
        _lib.count_launch()
        g.val = val
        return g

    def row_view(self, g: DeviceCSR, lo: int, hi: int) -> DeviceCSR:
        """Rows [lo, hi) of g as a CSR of its own (rowptr slice; colidx/val shared, offsets stay absolute)."""
        v = DeviceCSR(hi - lo, g.n_cols, g.rowptr[lo:hi + 1], g.colidx, g.val, chunk=DEFAULT_CHUNK)
        v.nnz = int(g.rowptr[hi]) - int(g.rowptr[lo])   # entries of the view (the arrays themselves are shared)
        v._struct = None
        return v

    def spmm(self, g, X, Y=None, resid=None, acc_in=None, acc_out=None, acc_div=1.0):
        g.spmm(X, Y=Y, resid=resid, acc_in=acc_in, acc_out=acc_out, acc_div=acc_div, want_y=Y is not None)

    def spmm_split(self, g, X, split_row, y_tail, Y=None, resid=None, acc_in=None, acc_out=None, acc_div=1.0):
        """One launch over ALL local rows: rows < split_row get the fused epilogue, rows >= split_row (the partial item
        rows) are stored raw to y_tail."""
        g.spmm(X, Y=Y, resid=resid, acc_in=acc_in, acc_out=acc_out, acc_div=acc_div, want_y=Y is not None,
               split_row=split_row, y_tail=y_tail)

    def accumulate(self, y, acc, resid, div, out):
        with torch.cuda.device(self.device):
            check(_lib.load().lgb_accumulate(ptr(y), ptr(acc), ptr(resid), y.numel(), float(div), ptr(out), stream()), "accumulate")
        _lib.count_launch()

    def zero(self, t):
        with torch.cuda.device(self.device):
            check(_lib.load().lgb_zero(ptr(t), t.numel() * t.element_size(), stream()), "zero")

    def bpr(self, Ef, E0, Ug, u, p, n, lam, B_norm, user_lo=0, user_hi=0, loss=None, dEf=None, dE0_users=None,
            dE0_items=None, gscale=1.0):
        """BPR over triples (u, p, n).  Either the caller already compacted them to this rank's users (local user
        ids, user_hi == 0), or it passes the GLOBAL batch with global user ids and the owned range
        [user_lo, user_hi): the kernel then skips foreign triples (static shapes, no host sync).  B_norm = global B."""
        d = Ef.shape[1]
        off = Ug * d * 4
        a = LgbBprArgs()
        a.uf, a.pf, a.nf = Ef.data_ptr(), Ef.data_ptr() + off, Ef.data_ptr() + off
        a.u0, a.p0, a.n0 = E0.data_ptr(), E0.data_ptr() + off, E0.data_ptr() + off
        a.iu, a.ip, a.in_ = ptr(u), ptr(p), ptr(n)
        a.B, a.B_norm, a.d, a.lambda_, a.gscale = u.numel(), int(B_norm), d, float(lam), float(gscale)
        a.user_lo, a.user_hi = int(user_lo), int(user_hi)
        if dEf is not None:
            a.duf, a.dpf, a.dnf = dEf.data_ptr(), dEf.data_ptr() + off, dEf.data_ptr() + off
        if dE0_users is not None:
            a.du0 = dE0_users.data_ptr()
        if dE0_items is not None:
            a.dp0 = a.dn0 = dE0_items.data_ptr()
        if loss is not None:
            if self._bpr_ws is None or self._bpr_ws.numel() < 2 * int(_lib.load().lgb_bpr_blocks(max(u.numel(), 1))):
                self._bpr_ws = _bpr_ws(max(u.numel(), 1), Ef.device)
            a.loss, a.ws = loss.data_ptr(), ptr(self._bpr_ws)
        if u.numel() == 0 or (user_hi > 0 and user_hi <= user_lo):      # nothing to do on this rank
            if loss is not None:
                self.zero(loss)
            return
        _bpr_launch(a, Ef.device)

    def adam(self, p, g, m, v, lr, beta1, beta2, eps, step):
        with torch.cuda.device(self.device):
            check(_lib.load().lgb_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), lr, beta1, beta2, eps, int(step), stream()),
                  "adam_step")
        _lib.count_launch()

    # collectives -------------------------------------------------------------------------------
    def all_reduce_async(self, t: torch.Tensor):
        """Sum-all-reduce of ``t`` on the comm stream, ordered after the work already queued on the
        compute stream; returns a handle whose wait() makes the compute stream wait for the result."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return _Done()
        cur = torch.cuda.current_stream(self.device)
        self.comm.wait_stream(cur)
        with torch.cuda.stream(self.comm):
            if self.exchange_events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(self.comm)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            if self.exchange_events is not None:
                e1.record(self.comm)
                self.exchange_events.append((e0, e1, t.numel() * t.element_size()))
        return _StreamWait(self.comm, cur)

    def all_reduce(self, t: torch.Tensor):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)


class SymmOps(CudaOps):
    """Item-block exchange WITHOUT NCCL: the block is staged in CUDA symmetric memory and summed by
    lgb_multimem_allreduce_f32 (NVSwitch multicast: multimem.ld_reduce + multimem.st, rank g owns slice g) or, when
    the fabric has no multicast, lgb_peer_allreduce_f32 (P2P loads/stores over NVLink), bracketed by symmetric-memory
    barriers.  Opt-in: ShardedLightGCN(..., exchange="symm").  Written in round 1 after the GPU budget was spent:
    compiles (SASS shows LDGMC.E.ADD.F32x4), logic mirrors torch's two-shot multimem all-reduce, NOT yet run on B200."""

    def __init__(self, device: torch.device, group=None):
        super().__init__(device, group)
        import torch.distributed._symmetric_memory as symm_mem
        self._symm = symm_mem
        self._stage = {}

    def _staging(self, numel: int):
        if numel not in self._stage:
            t = self._symm.empty(numel, dtype=torch.float32, device=self.device)
            h = self._symm.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
            self._stage[numel] = (t, h)
        return self._stage[numel]

    def all_reduce_async(self, t: torch.Tensor):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return _Done()
        n = t.numel()
        if n % 4 != 0 or not t.is_contiguous():
            return super().all_reduce_async(t)
        stage, h = self._staging(n)
        lib = _lib.load()
        cur = torch.cuda.current_stream(self.device)
        self.comm.wait_stream(cur)
        with torch.cuda.stream(self.comm):
            stage.copy_(t.reshape(-1))
            h.barrier(channel=0, timeout_ms=20000)                                   # every rank's partial sums are staged
            off = int(getattr(h, "offset", 0))
            mc = int(h.multicast_ptr) if h.multicast_ptr else 0
            with torch.cuda.device(self.device):
                if mc:
                    check(lib.lgb_multimem_allreduce_f32(mc + off, n, h.rank, h.world_size, self.comm.cuda_stream),
                          "multimem_allreduce")
                else:
                    import ctypes as C
                    arr = (C.c_uint64 * h.world_size)(*[int(p) + off for p in h.buffer_ptrs])
                    check(lib.lgb_peer_allreduce_f32(arr, n, h.rank, h.world_size, self.comm.cuda_stream), "peer_allreduce")
            _lib.count_launch()
            h.barrier(channel=1, timeout_ms=20000)                                   # every slice is republished on every rank
            t.reshape(-1).copy_(stage)
        return _StreamWait(self.comm, cur)


class _Done:
    def wait(self):
        pass


class _StreamWait:
    def __init__(self, comm, cur):
        self.comm, self.cur = comm, cur

    def wait(self):
        self.cur.wait_stream(self.comm)


# --------------------------------------------------------------------------------------------
# the sharded engine
# --------------------------------------------------------------------------------------------
class ShardedLightGCN:
    """LightGCN fwd + BPR + bwd over a user-sharded graph.  Every rank passes the same (users, items) COO
    (or any superset of its own users' edges plus all item-degree information via ``item_degree``)."""

    def __init__(self, num_users: int, num_items: int, embedding_dim: int, num_iterations: int,
                 users: torch.Tensor, items: torch.Tensor, device, group=None, ops=None,
                 rank: Optional[int] = None, world: Optional[int] = None, init_tables=None,
                 schedule: str = "layer", static_batch: bool = False, exchange: str = "nccl"):
        self.U, self.I, self.d, self.K = int(num_users), int(num_items), int(embedding_dim), int(num_iterations)
        self.device = torch.device(device)
        if schedule not in ("layer", "pipelined", "merged"):
            raise ValueError(f"schedule={schedule!r}")
        self.schedule, self.static_batch = schedule, bool(static_batch)
        inited = dist.is_available() and dist.is_initialized()
        self.rank = rank if rank is not None else (dist.get_rank(group) if inited else 0)
        self.world = world if world is not None else (dist.get_world_size(group) if inited else 1)
        if exchange not in ("nccl", "symm"):
            raise ValueError(f"exchange={exchange!r}")
        self.ops = ops if ops is not None else (SymmOps if exchange == "symm" else CudaOps)(self.device, group)
        users, items = users.to(self.device), items.to(self.device)

        udeg = torch.bincount(users, minlength=self.U)
        self.bounds = balanced_user_bounds(udeg, self.world)
        self.lo, self.hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.Ug = self.hi - self.lo
        row, col, lu, li = local_block(users, items, self.lo, self.hi)
        self.local_edges = int(lu.numel())
        # global degrees: users are wholly local; item degrees are summed over ranks
        ideg = torch.bincount(li, minlength=self.I).to(torch.float32)
        self.ops.all_reduce(ideg)
        deg = torch.cat([udeg[self.lo:self.hi].to(torch.float32), ideg])
        dinv = torch.where(deg > 0, 1.0 / torch.sqrt(deg), torch.zeros_like(deg))
        n = self.Ug + self.I
        self.n = n
        g = self.ops.build_graph(row, col, n, dinv)
        self.g_users = self.ops.row_view(g, 0, self.Ug)
        self.g_items = self.ops.row_view(g, self.Ug, n)
        self.g_full = g
        self.g_all = self.ops.row_view(g, 0, n) if schedule == "merged" else None   # all rows + split plan

        # parameters: one local table; the item block is identical on every rank
        f32 = dict(dtype=torch.float32, device=self.device)
        self.table = torch.empty(n, self.d, **f32)
        if init_tables is not None:
            Wu, Wi = init_tables
            self.table[: self.Ug].copy_(Wu[self.lo:self.hi])
            self.table[self.Ug:].copy_(Wi)
        else:
            gen = torch.Generator(device=self.device).manual_seed(1000 + self.rank)
            self.table[: self.Ug].normal_(0, 0.1, generator=gen)
            gen_i = torch.Generator(device=self.device).manual_seed(999)
            self.table[self.Ug:].normal_(0, 0.1, generator=gen_i)
        self.grad = torch.empty(n, self.d, **f32)
        self.E_f = torch.empty(n, self.d, **f32)
        self._ya = torch.empty(n, self.d, **f32)
        self._yb = torch.empty(n, self.d, **f32)
        self._r = torch.empty(n + 1, self.d, **f32)   # + one row that carries the scalar loss through the all-reduce
        self._loss = torch.zeros((), **f32)
        self.loss = self._loss

    def graphs(self):
        return [self.g_users, self.g_items]

    def autotune(self):
        """Per-rank plan-time choice of the SpMM kernel variant for each row view (DeviceCSR.autotune); ranks may choose
        differently -- the item rows are all-reduced, so the replicated blocks stay identical."""
        chunks = (DEFAULT_CHUNK, 256)      # a rank's launch is bounded below by one slice's serial chain: try shorter slices
        views = {"users": (self.g_users, True), "items": (self.g_items, False)}
        if self.g_all is not None:
            views["all"] = (self.g_all, True)
        out = {}
        for name, (g, fused) in views.items():
            g.autotune(self.d, fused_epilogue=fused, chunks=chunks)
            out[name] = dict(g.autotune_report.get("chosen", {"variant": g.variant}), ms=g.autotune_report["ms"],
                             rejected=g.autotune_report["rejected"])
        return out

    # ---- optimizer step on the local shard (run_pipeline_lightgcn.py:103,159 -- optim.Adam over both tables) ------------
    def adam_step(self, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8) -> None:
        """One fused Adam step (lgb_adam_step, same arithmetic as torch.optim.Adam) on the local table with the gradients the
        last ``fused_step`` left in ``self.grad``: the owned user rows are updated by their owner only; the item rows carry
        identical gradients on every rank (they were all-reduced), so the replicated item block stays bit-identical without
        a broadcast.  State (exp_avg, exp_avg_sq, step) lives on the engine."""
        if not hasattr(self, "_adam"):
            self._adam = dict(step=0, m=torch.zeros_like(self.table), v=torch.zeros_like(self.table))
        st = self._adam
        st["step"] += 1
        self.ops.adam(self.table, self.grad, st["m"], st["v"], float(lr), float(betas[0]), float(betas[1]), float(eps), st["step"])

    def autotune_step(self, user_indices, pos_item_indices, neg_item_indices, lambda_val: float, reps: int = 5,
                      candidates=((None, False), (None, True)), timer=None) -> dict:
        """Plan-time choice of the step form on THIS machine: every candidate (schedule, static_batch) -- by default the
        measured host-filtered step and the static-shape step (kernel-side owned-user filter, loss riding on the gradient
        all-reduce, that all-reduce hidden behind the first backward SpMM) under the current schedule -- must reproduce the
        first candidate's loss and gradients on the given batch on EVERY rank, and is then timed (barrier, CUDA events, max
        over ranks); the fastest is kept.  Every rank runs the same candidates in the same order, so the collective sequence
        stays matched.  Returns a report."""
        inited = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.ops.group) > 1

        def agree(value: float, op) -> float:
            t = torch.tensor([value], dtype=torch.float64, device=self.device)
            if inited:
                dist.all_reduce(t, op=op, group=self.ops.group)
            return float(t)

        def default_timer(fn):
            if inited:
                dist.barrier(group=self.ops.group)
            torch.cuda.synchronize(self.device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            e1.synchronize()
            return e0.elapsed_time(e1) / reps
        timer = timer or default_timer
        keep = (self.schedule, self.static_batch)
        report, ref, best, best_ms = {"ms": {}, "rejected": {}}, None, keep, float("inf")
        for schedule, static in candidates:
            schedule = schedule or keep[0]
            if schedule == "merged" and self.g_all is None:      # the all-rows view (+ its plan and kernel choice) is built on demand
                self.g_all = self.ops.row_view(self.g_full, 0, self.n)
                if self.g_users.variant is not None:
                    self.g_all.autotune(self.d, chunks=(DEFAULT_CHUNK, 256))
            key = f"{schedule}/{'static' if static else 'host-filtered'}"
            self.schedule, self.static_batch = schedule, bool(static)
            loss = self.fused_step(user_indices, pos_item_indices, neg_item_indices, lambda_val).clone()
            grad = self.grad.clone()
            if ref is None:
                ref = (loss, grad)
            else:
                same = bool(torch.allclose(loss, ref[0], rtol=1e-5, atol=1e-7)) and \
                    bool(torch.allclose(grad, ref[1], rtol=1e-4, atol=1e-6 * float(ref[1].abs().max()) + 1e-12))
                if agree(1.0 if same else 0.0, dist.ReduceOp.MIN) < 0.5:
                    report["rejected"][key] = "loss / gradients differ from the first candidate on at least one rank"
                    continue
            ms = agree(timer(lambda: self.fused_step(user_indices, pos_item_indices, neg_item_indices, lambda_val)), dist.ReduceOp.MAX)
            report["ms"][key] = ms
            if ms < best_ms:
                best, best_ms = (schedule, bool(static)), ms
        self.schedule, self.static_batch = best
        report["chosen"] = {"schedule": best[0], "static_batch": best[1]}
        return report

    @property
    def users_weight(self):
        return self.table[: self.Ug]

    @property
    def items_weight(self):
        return self.table[self.Ug:]

    # ---- schedule "layer" (default; measured on 2/4/8 B200): one all-reduce per layer, overlapped with the users SpMM ----
    def _layer(self, X, Y, resid=None, acc_in=None, acc_out=None, acc_div=1.0, write_y=True):
        """Y[:Ug] = G_users X (+resid) ; Y[Ug:] = allreduce(G_items X) (+resid);
        acc_out (optional) = (acc_in + that) / acc_div on both row blocks."""
        Ug, ops = self.Ug, self.ops
        yi = Y[Ug:]
        ops.spmm(self.g_items, X, Y=yi)                                     # partial item rows
        h = ops.all_reduce_async(yi)                                          # ... summed over ranks (comm stream)
        ops.spmm(self.g_users, X, Y=Y[:Ug] if write_y else None,              # overlaps with the all-reduce
                 resid=None if resid is None else resid[:Ug],
                 acc_in=None if acc_in is None else acc_in[:Ug],
                 acc_out=None if acc_out is None else acc_out[:Ug], acc_div=acc_div)
        self._finish_items(h, Y, resid, acc_in, acc_out, acc_div)

    def _finish_items(self, h, y, resid, acc_in, acc_out, div):
        """Wait for the all-reduce of y's item rows, then apply the epilogue the SpMM could not fuse for them."""
        Ug, ops = self.Ug, self.ops
        h.wait()
        yi = y[Ug:]
        if acc_out is not None:
            ops.accumulate(yi, None if acc_in is None else acc_in[Ug:], None if resid is None else resid[Ug:], div, acc_out[Ug:])
        elif resid is not None:
            ops.accumulate(yi, None, resid[Ug:], 1.0, yi)

    # ---- schedule "pipelined" (opt-in): the all-reduce hidden behind TWO SpMM launches -----------------------------
    # Bipartite dependencies:  users^{k+1} <- items^k (needs the all-reduced item rows)
    #                          items^{k+1} <- users^k (local rows only)
    # so all-reduce(items^{k+1}) is only needed by users^{k+2}: it can overlap with users^{k+1} AND items^{k+2}.
    # Validated against the oracle over gloo (tests/test_dist_gloo.py); B200 measurement pending.
    def _propagate_pipelined(self, x0, K, resid=None, acc0=None, acc=None, acc_div_last=1.0, out_last=None,
                             before_last_reduce=None, x0_items_pending=None):
        Ug, ops = self.Ug, self.ops
        bufs = [self._ya, self._yb]
        x, pending = x0, None            # pending = epilogue of the layer whose item rows are still being reduced
        for k in range(K):
            last = k == K - 1
            y = out_last if (last and out_last is not None) else bufs[k % 2]
            ops.spmm(self.g_items, x, Y=y[Ug:])                               # items^{k+1} partial (reads user rows of x)
            if k == 0 and x0_items_pending is not None:                        # x0's item rows were still being all-reduced
                x0_items_pending.wait()
            if last and before_last_reduce is not None:
                before_last_reduce(y[Ug:])
            if pending is not None:                                            # x's item rows must be complete now
                self._finish_items(*pending)
            h = ops.all_reduce_async(y[Ug:])
            div = acc_div_last if last else 1.0
            a_in = None if acc is None else (acc0 if k == 0 else acc)
            ops.spmm(self.g_users, x, Y=None if (last and acc is not None) else y[:Ug],
                     resid=None if resid is None else resid[:Ug],
                     acc_in=None if a_in is None else a_in[:Ug],
                     acc_out=None if acc is None else acc[:Ug], acc_div=div)
            pending = (h, y, resid, a_in, acc, div)
            x = y
        if pending is not None:
            self._finish_items(*pending)
        return x

    # ---- schedule "merged" (opt-in): ONE launch per layer over all local rows (split epilogue), exchange exposed ------
    # Halves the launch count and doubles the work per launch (at 1/8 of the graph a launch is bound by per-row latency
    # chains, not bytes); pays with a fully exposed exchange, so it is meant for the fast multimem exchange.
    # Validated over gloo; B200 measurement pending.
    def _layer_merged(self, X, Y, resid=None, acc_in=None, acc_out=None, acc_div=1.0, write_y=True, before_reduce=None):
        Ug, ops = self.Ug, self.ops
        ops.spmm_split(self.g_all, X, Ug, Y[Ug:], Y=Y if write_y else None, resid=resid, acc_in=acc_in, acc_out=acc_out,
                       acc_div=acc_div)
        if before_reduce is not None:
            before_reduce(Y[Ug:])
        self._finish_items(ops.all_reduce_async(Y[Ug:]), Y, resid, acc_in, acc_out, acc_div)

    def forward(self) -> torch.Tensor:
        """E_f = mean_k A^k E0 on the local rows (item rows replicated)."""
        K, E0, Ef = self.K, self.table, self.E_f
        if K == 0:
            Ef.copy_(E0)
            return Ef
        if self.schedule == "merged":
            x, y = E0, self._ya
            for k in range(K):
                last = k == K - 1
                self._layer_merged(x, y, acc_in=E0 if k == 0 else Ef, acc_out=Ef, acc_div=float(K + 1) if last else 1.0,
                                   write_y=not last)
                x, y = y, (self._yb if y is self._ya else self._ya)
            return Ef
        if self.schedule == "pipelined":
            self._propagate_pipelined(E0, K, acc0=E0, acc=Ef, acc_div_last=float(K + 1))
            return Ef
        x, y = E0, self._ya
        for k in range(K):
            last = k == K - 1
            self._layer(x, y, acc_in=E0 if k == 0 else Ef, acc_out=Ef, acc_div=float(K + 1) if last else 1.0,
                        write_y=not last)
            x, y = y, (self._yb if y is self._ya else self._ya)
        return Ef

    def backward(self, r: torch.Tensor, before_last_reduce=None, r_items_pending=None) -> torch.Tensor:
        """grad = sum_k (A^T)^k r with r = dE_f/(K+1).  The local block is symmetric, so A^T is the same pair of row views.
        r's item rows must be the sum over ranks: either already (r_items_pending None) or once r_items_pending.wait()
        returns -- the first layer's items SpMM only reads r's USER rows, so that all-reduce is hidden behind it."""
        K = self.K
        if K == 0 or self.schedule == "merged":
            if r_items_pending is not None:          # nothing to hide it behind
                r_items_pending.wait()
                r_items_pending = None
        if K == 0:
            self.grad.copy_(r)
            return self.grad
        if self.schedule == "pipelined":
            return self._propagate_pipelined(r, K, resid=r, out_last=self.grad, before_last_reduce=before_last_reduce,
                                             x0_items_pending=r_items_pending)
        g = r
        bufs = [self._ya, self._yb]
        for k in range(K):
            last = k == K - 1
            dst = self.grad if last else bufs[k % 2]
            if self.schedule == "merged":
                self._layer_merged(g, dst, resid=r, before_reduce=before_last_reduce if last else None)
                g = dst
                continue
            if (last and before_last_reduce is not None) or (k == 0 and r_items_pending is not None):
                Ug, ops = self.Ug, self.ops
                ops.spmm(self.g_items, g, Y=dst[Ug:])                     # reads user rows of g only
                if k == 0 and r_items_pending is not None:
                    r_items_pending.wait()                                # from here on r's item rows are needed
                if last and before_last_reduce is not None:
                    before_last_reduce(dst[Ug:])                          # fold extra partial item-row terms into this all-reduce
                h = ops.all_reduce_async(dst[Ug:])
                ops.spmm(self.g_users, g, Y=dst[:Ug], resid=r[:Ug])
                self._finish_items(h, dst, r, None, None, 1.0)
            else:
                self._layer(g, dst, resid=r)
            g = dst
        return self.grad

    @torch.no_grad()
    def fused_step(self, user_indices: torch.Tensor, pos_item_indices: torch.Tensor, neg_item_indices: torch.Tensor,
                   lambda_val: float) -> torch.Tensor:
        """Global batch in (the same B triples on every rank), global loss out (0-dim tensor, identical on every
        rank); gradients in ``self.grad`` (rows [:Ug] owned users, rows [Ug:] replicated items, identical on every
        rank)."""
        if self.static_batch:
            return self._fused_step_static(user_indices, pos_item_indices, neg_item_indices, lambda_val)
        ops, Ug, K = self.ops, self.Ug, self.K
        u, p, n = (t.to(self.device) for t in (user_indices, pos_item_indices, neg_item_indices))
        B = u.numel()
        mine = (u >= self.lo) & (u < self.hi)                               # triples of the users this rank owns
        lu, lp, ln = (u[mine] - self.lo).contiguous(), p[mine].contiguous(), n[mine].contiguous()
        Ef = self.forward()
        r = self._r[: self.n]
        ops.zero(r)
        loss = self._loss
        ops.bpr(Ef, self.table, Ug, lu, lp, ln, lambda_val, B, loss=loss, dEf=r, gscale=1.0 / (K + 1))
        # item-row gradient contributions of the local triples are partial sums: reduce them (and the loss)
        h = ops.all_reduce_async(r[Ug:])
        ops.all_reduce(loss)
        h.wait()
        G = self.backward(r)
        # + 2*lambda*E0 on the batch rows: users locally; items are partial over ranks -> reduce a small buffer
        reg_items = self._ya[Ug:]
        ops.zero(reg_items)
        ops.bpr(Ef, self.table, Ug, lu, lp, ln, lambda_val, B, dE0_users=G, dE0_items=reg_items)
        ops.all_reduce_async(reg_items).wait()
        ops.accumulate(reg_items, G[Ug:], None, 1.0, G[Ug:])
        self.loss = loss
        return loss

    @torch.no_grad()
    def _fused_step_static(self, user_indices, pos_item_indices, neg_item_indices, lambda_val: float) -> torch.Tensor:
        """Same step with static shapes and no host synchronisation (CUDA-graph capturable): the kernel filters the
        global batch by the owned user range, the loss rides in an extra row of the residual buffer (one all-reduce
        for item-row gradients + loss) and the item regulariser terms ride on the last backward all-reduce.
        Validated over gloo (tests/test_dist_gloo.py); B200 measurement pending."""
        ops, Ug, K = self.ops, self.Ug, self.K
        u, p, n = (_lib.i64c(t.to(self.device)) for t in (user_indices, pos_item_indices, neg_item_indices))
        B = u.numel()
        Ef = self.forward()
        r = self._r                                   # [n + 1, d]: the extra row carries the loss through the all-reduce
        ops.zero(r)
        loss_slot = r[self.n, :1].view(())
        flt = dict(user_lo=self.lo, user_hi=self.hi)
        ops.bpr(Ef, self.table, Ug, u, p, n, lambda_val, B, loss=loss_slot, dEf=r, gscale=1.0 / (K + 1), **flt)
        h_r = ops.all_reduce_async(r[Ug:])            # item-row gradients + loss: ONE all-reduce, hidden behind the first
        rr = r[: self.n]                              # backward layer's items SpMM (which only reads r's user rows)

        def add_item_reg(items_partial):              # 2*lambda*E0[p], E0[n] of the local triples
            ops.bpr(Ef, self.table, Ug, u, p, n, lambda_val, B, dE0_items=items_partial, **flt)
        if K == 0:
            G = self.backward(rr, r_items_pending=h_r)
            reg_items = self._ya[Ug:]
            ops.zero(reg_items)
            add_item_reg(reg_items)
            ops.all_reduce_async(reg_items).wait()
            ops.accumulate(reg_items, G[Ug:], None, 1.0, G[Ug:])
        else:
            G = self.backward(rr, before_last_reduce=add_item_reg, r_items_pending=h_r)
        ops.bpr(Ef, self.table, Ug, u, p, n, lambda_val, B, dE0_users=G, **flt)   # owned users: local
        self.loss = loss_slot
        return loss_slot

    # ---- CUDA-graph replay of the whole step (opt-in; the 8-GPU step is ~40 launches of 0.1-0.2 ms each) -------------
    def capture(self, batch_size: int, lambda_val: float):
        """Capture the static-shape step for a fixed batch size; returns step(u, p, n) -> loss that replays the graph.
        Opt-in and not yet measured on B200 (bench.py --graph)."""
        dev = self.device
        self.static_batch = True
        su, sp, sn = (torch.zeros(batch_size, dtype=torch.int64, device=dev) for _ in range(3))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self.fused_step(su, sp, sn, lambda_val)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = self.fused_step(su, sp, sn, lambda_val)

        def step(u, p, n):
            su.copy_(u, non_blocking=True); sp.copy_(p, non_blocking=True); sn.copy_(n, non_blocking=True)
            graph.replay()
            return loss
        self._graph = graph
        return step
