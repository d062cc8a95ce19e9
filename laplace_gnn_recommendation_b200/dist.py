"""Multi-GPU LightGCN: users range-sharded, items replicated, one exchange of the [I, d] item block per propagation
layer (SURVEY.md 8e "bipartite-aware 1.5-D row sharding").  One process per GPU; ``torch.distributed`` is the plumbing
(rendezvous, symmetric-memory allocation), the exchange itself is a hand-written kernel over NVSwitch multicast /
NVLink peer memory (csrc/peer.cu), NCCL all-reduce being the fallback.  The reference has no distributed code at all:
this layer is new.

Per rank g (owning users [u_g, u_{g+1}), split at nnz-balanced points):
    table     [U_g + I, d]     rows 0..U_g-1 = owned user embeddings, rows U_g.. = ALL item embeddings (replicated)
    G_users   [U_g, I]   CSR   R_g    with the GLOBAL symmetric normalisation (item degrees all-reduced once at build time)
    G_items   [I, U_g]   CSR   R_g^T  (partial sums over the owned users only)
A propagation layer k maps (xu, xi) -> (yu, yi):
    yu = G_users xi                          complete locally (items are replicated)
    yi = exchange(G_items xu)                sum over ranks of the partial item rows
yu_k only needs xi_k and yi_k only needs xu_k, so the K layers fall apart into TWO INDEPENDENT CHAINS
    chain 0:  U(0) -> I(1) -> X(1) -> U(2) -> I(3) -> X(3) ...          (U = users SpMM, I = items SpMM, X = exchange)
    chain 1:  I(0) -> X(0) -> U(1) -> I(2) -> X(2) -> U(3) ...
which run on two streams: while one chain waits for its exchange the other computes, and the two SpMM kernels of a
layer fill each other's tail.  The exchange kernels run on high-priority streams so that their few CTAs are scheduled
ahead of the queued SpMM CTAs.  The layer mean is one pass over the kept layer outputs at the end (lgb_mean_rows, same
left-to-right order as the single-GPU fused epilogue).  The backward is the same pair of chains on the (symmetric)
block with the residual r = dE_f/(K+1): user rows add r in the SpMM epilogue, the item-row residual is added to rank
0's partial sums so that the exchange delivers it exactly once.

BPR: the batch's user rows (final and layer-0: 2*B*d floats) are gathered by their owners into a small symmetric
buffer and exchanged; every rank then evaluates ALL B triples -- loss and item-row gradients come out complete and
bit-identical on every rank without a 27 MB collective -- and scatters the user-row gradients of the users it owns.
A step therefore moves 2K item blocks plus one 2*B*d block; user embeddings never move.

The kernel / collective calls go through a small ``ops`` object so that the partition + exchange logic can be exercised
on the CPU with gloo and an oracle-backed ops object in tests/ (the product ops objects are CUDA-only and fail loudly
otherwise).
"""
from __future__ import annotations

import contextlib
import os
import ctypes as C
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from ._lib import LgbBprArgs, LgbExchange, check, ptr, stream
from .bpr import _launch as _bpr_launch, _ws as _bpr_ws
from .csr import DEFAULT_CHUNK, DeviceCSR, rows_bitmap

N_CHANNELS = 3          # exchange channels: one per layer chain + the BPR batch rows


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
# CUDA ops objects
# --------------------------------------------------------------------------------------------
class _Done:
    def wait(self):
        pass


class _StreamWait:
    """Handle of an exchange running on a side stream: wait() makes the stream that is current THEN wait for it."""

    def __init__(self, comm, device):
        self.comm, self.device = comm, device

    def wait(self):
        torch.cuda.current_stream(self.device).wait_stream(self.comm)


class CudaOps:
    """liblaplace_b200 kernels; the item-block exchange is an NCCL all-reduce on a high-priority side stream (the fallback
    of ``SymmOps``, and the only collective path on a fabric without peer access)."""

    kind = "nccl"

    def __init__(self, device: torch.device, group=None):
        if device.type != "cuda":
            raise RuntimeError("ShardedLightGCN needs CUDA devices (no CPU fallback); tests inject their own ops object")
        self.device, self.group = device, group
        hi_pri = torch.cuda.Stream.priority_range()[1] if hasattr(torch.cuda.Stream, "priority_range") else -1
        # one comm stream per exchange channel; NCCL serialises its collectives anyway, so this class uses one for all
        self.comm = [torch.cuda.Stream(device=device, priority=hi_pri)] * N_CHANNELS
        self.chains = [torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)]
        self._bpr_ws = None
        self.exchange_events = None     # set to a list to record (start, end, bytes) CUDA events of every item-block exchange

    # streams --------------------------------------------------------------------------------------
    def fork(self, two_streams: bool):
        """Stream contexts of the two layer chains, ordered after the work queued on the current stream."""
        if not two_streams:
            return [contextlib.nullcontext(), contextlib.nullcontext()]
        cur = torch.cuda.current_stream(self.device)
        for s in self.chains:
            s.wait_stream(cur)
        return [torch.cuda.stream(s) for s in self.chains]

    def join(self, two_streams: bool):
        if two_streams:
            cur = torch.cuda.current_stream(self.device)
            for s in self.chains:
                cur.wait_stream(s)

    # graph ----------------------------------------------------------------------------------------
    def build_views(self, row, col, Ug: int, I: int, dinv):
        """The local symmetric block [[0, R_g], [R_g^T, 0]] -> (G_users [Ug, I], G_items [I, Ug]) sharing one colidx / val array
        pair: rows [0, Ug) with their column ids shifted to item numbering, rows [Ug, Ug+I) as they are (user numbering)."""
        n = Ug + I
        g = DeviceCSR.from_coo(row, col, n, n, chunk=0)
        val = torch.empty(g.nnz, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(_lib.load().lgb_gcn_values(ptr(g.rowptr), ptr(g.colidx), n, g.nnz, ptr(dinv), ptr(val), stream()), "gcn_values")
        _lib.count_launch()
        split = int(g.rowptr[Ug])                       # entries of the user rows come first (CSR order)
        g.colidx[:split] -= Ug
        gu = DeviceCSR(Ug, I, g.rowptr[: Ug + 1], g.colidx, val, chunk=DEFAULT_CHUNK)
        gi = DeviceCSR(I, Ug, g.rowptr[Ug:], g.colidx, val, chunk=DEFAULT_CHUNK)
        gu.nnz, gi.nnz = split, g.nnz - split           # entries of the view (the arrays themselves are shared)
        gu._struct = gi._struct = None
        return gu, gi

    def alloc_exchange(self, shapes):
        """Buffers that ``exchange_async`` can sum over ranks in place."""
        return [torch.zeros(*s, dtype=torch.float32, device=self.device) for s in shapes]

    # kernels --------------------------------------------------------------------------------------
    def spmm(self, g, X, Y, resid=None, x_rows=None, resid_rows=None):
        if g.n_rows == 0:
            return
        if g.n_cols == 0:                       # a rank without users (world > #users): its partial sums are zero
            Y.zero_() if resid is None else Y.copy_(resid)
            return
        g.spmm(X, Y=Y, resid=resid, x_rows=x_rows, resid_rows=resid_rows if (x_rows is not None and resid is not None) else None)

    def rows_bitmap(self, n, index_lists, key):
        """Bitmap of the rows idx + offset (``(idx, offset)`` pairs; ids outside [0, n) ignored) for ``spmm(x_rows=)``; one
        persistent buffer per ``key`` (graph capture)."""
        if n <= 0:
            return None
        bufs = self.__dict__.setdefault("_bitmaps", {})
        buf = bufs.get(key)
        if buf is None or buf.numel() * 32 < n:
            buf = bufs[key] = torch.empty((n + 31) // 32, dtype=torch.int32, device=self.device)
        return rows_bitmap(n, index_lists, self.device, out=buf)

    def row_view(self, g: DeviceCSR, lo: int, hi: int) -> DeviceCSR:
        """Rows [lo, hi) of g as a CSR of its own (rowptr slice; colidx/val shared, offsets stay absolute) -- probes."""
        v = DeviceCSR(hi - lo, g.n_cols, g.rowptr[lo:hi + 1], g.colidx, g.val, chunk=DEFAULT_CHUNK)
        v.nnz = int(g.rowptr[hi]) - int(g.rowptr[lo])
        v._struct = None
        return v

    def mean_rows(self, srcs, div, out):
        arr = (C.c_void_p * len(srcs))(*[t.data_ptr() for t in srcs])
        with torch.cuda.device(self.device):
            check(_lib.load().lgb_mean_rows(arr, len(srcs), out.numel(), float(div), ptr(out), stream()), "mean_rows")
        _lib.count_launch()

    def accumulate(self, y, acc, resid, div, out):
        with torch.cuda.device(self.device):
            check(_lib.load().lgb_accumulate(ptr(y), ptr(acc), ptr(resid), y.numel(), float(div), ptr(out), stream()), "accumulate")
        _lib.count_launch()

    def zero(self, t):
        with torch.cuda.device(self.device):
            check(_lib.load().lgb_zero(ptr(t), t.numel() * t.element_size(), stream()), "zero")

    def gather_owned(self, src, idx, lo, hi, dst):
        with torch.cuda.device(self.device):
            check(_lib.load().lgb_gather_rows_owned(ptr(src), ptr(idx), idx.numel(), src.shape[1], int(lo), int(hi), ptr(dst), stream()),
                  "gather_rows_owned")
        _lib.count_launch()

    def scatter_add_owned(self, src, idx, lo, hi, dst):
        with torch.cuda.device(self.device):
            check(_lib.load().lgb_scatter_add_rows_owned(ptr(src), ptr(idx), idx.numel(), src.shape[1], int(lo), int(hi), ptr(dst), stream()),
                  "scatter_add_rows_owned")
        _lib.count_launch()

    def bpr(self, uf, u0, pf, p0, iu, ip, in_, lam, B_norm, gscale=1.0, user_lo=0, user_hi=0, user_rows_only=False,
            loss=None, duf=None, du0=None, dpf=None, dp0=None):
        """lgb_bpr on explicit operand tables: user operand (uf, u0) indexed by iu, positive AND negative operand (pf, p0) by
        ip / in_; gradients accumulate into duf / du0 (user side) and dpf / dp0 (both item sides)."""
        d = pf.shape[1]
        a = LgbBprArgs()
        a.uf, a.u0, a.pf, a.p0, a.nf, a.n0 = ptr(uf), ptr(u0), ptr(pf), ptr(p0), ptr(pf), ptr(p0)
        a.iu, a.ip, a.in_ = ptr(iu), ptr(ip), ptr(in_)
        a.B, a.B_norm, a.d, a.lambda_, a.gscale = iu.numel(), int(B_norm), d, float(lam), float(gscale)
        a.user_lo, a.user_hi, a.flags = int(user_lo), int(user_hi), 1 if user_rows_only else 0
        a.duf, a.du0 = ptr(duf), ptr(du0)
        a.dpf = a.dnf = ptr(dpf)
        a.dp0 = a.dn0 = ptr(dp0)
        if loss is not None:
            if self._bpr_ws is None or self._bpr_ws.numel() < 2 * int(_lib.load().lgb_bpr_blocks(max(iu.numel(), 1))):
                self._bpr_ws = _bpr_ws(max(iu.numel(), 1), pf.device)
            a.loss, a.ws = loss.data_ptr(), ptr(self._bpr_ws)
        if iu.numel() == 0:
            if loss is not None:
                self.zero(loss)
            return
        _bpr_launch(a, pf.device)

    def adam(self, p, g, m, v, lr, beta1, beta2, eps, step):
        with torch.cuda.device(self.device):
            check(_lib.load().lgb_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), lr, beta1, beta2, eps, int(step), stream()),
                  "adam_step")
        _lib.count_launch()

    # collectives ----------------------------------------------------------------------------------
    def _multi(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def exchange_async(self, t: torch.Tensor, channel: int = 0):
        """In-place sum over ranks of a buffer from ``alloc_exchange``, on the channel's high-priority side stream, ordered after
        the work already queued on the current stream; returns a handle whose wait() makes the then-current stream wait."""
        if not self._multi():
            return _Done()
        cur = torch.cuda.current_stream(self.device)
        comm = self.comm[channel]
        comm.wait_stream(cur)
        with torch.cuda.stream(comm):
            ev = self._ev_start(comm)
            self._exchange(t, channel, comm)
            self._ev_end(ev, comm, t)
        return _StreamWait(comm, self.device)

    all_reduce_async = exchange_async          # older name (dist_rows.py, tools)

    def _exchange(self, t, channel, comm):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def _ev_start(self, comm):
        if self.exchange_events is None:
            return None
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record(comm)
        return e0

    def _ev_end(self, e0, comm, t):
        if e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(comm)
            self.exchange_events.append((e0, e1, t.numel() * t.element_size()))

    def all_reduce(self, t: torch.Tensor):
        if self._multi():
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)


class SymmOps(CudaOps):
    """Item-block exchange WITHOUT NCCL: the exchange buffers live in ONE arena of CUDA symmetric memory (every rank maps every
    peer's copy and, with NVLS, one multicast address that fans out to all copies) and are summed in place by
    ``lgb_exchange_allreduce_f32`` -- entry barrier, multimem.ld_reduce of the rank's slice (reduced inside the NVSwitch),
    multimem.st back to all copies, exit barrier, all in one launch, no host synchronisation, CUDA-graph replayable.  Without
    multicast the same kernel uses peer loads / stores over NVLink."""

    kind = "symm"

    def __init__(self, device: torch.device, group=None):
        super().__init__(device, group)
        import torch.distributed._symmetric_memory as symm_mem
        self._symm = symm_mem
        hi_pri = torch.cuda.Stream.priority_range()[1] if hasattr(torch.cuda.Stream, "priority_range") else -1
        self.comm = [torch.cuda.Stream(device=device, priority=hi_pri) for _ in range(N_CHANNELS)]
        self._arena = self._x = self._handle = None
        self.multicast, self.mode, self.flags, self.mode_report = False, None, 0, None

    def alloc_exchange(self, shapes):
        if not self._multi():
            return super().alloc_exchange(shapes)
        if self._arena is not None:
            raise RuntimeError("SymmOps.alloc_exchange: one arena per ops object (allocate every exchange buffer in one call)")
        lib = _lib.load()
        grp = self.group if self.group is not None else dist.group.WORLD
        world, rank = dist.get_world_size(grp), dist.get_rank(grp)
        sizes = [int(torch.Size(s).numel()) for s in shapes]
        offs, total = [], 0
        for n in sizes:
            offs.append(total)
            total += (n + 63) // 64 * 64                       # 256-byte aligned sub-buffers
        pad_words = N_CHANNELS * int(lib.lgb_exchange_pad_words(world))
        pad_off = total
        total += (pad_words + 63) // 64 * 64
        arena = self._symm.empty(total, dtype=torch.float32, device=self.device)
        arena.zero_()                                          # data AND signal slots start at 0
        torch.cuda.synchronize(self.device)
        h = self._symm.rendezvous(arena, grp)
        base_off = int(getattr(h, "offset", 0) or 0)
        x = LgbExchange()
        mc = int(h.multicast_ptr) if getattr(h, "multicast_ptr", 0) else 0
        x.multicast_base = (mc + base_off) if mc else None
        for r in range(world):
            x.peer_base[r] = int(h.buffer_ptrs[r]) + base_off
            x.pad_base[r] = int(h.buffer_ptrs[r]) + base_off + pad_off * 4
        x.rank, x.world, x.n_channels = rank, world, N_CHANNELS
        self._arena, self._x, self._handle, self.multicast = arena, x, h, bool(mc)
        dist.barrier(group=grp)                                # every rank has zeroed its arena before anybody signals
        bufs = [arena[o:o + n].view(*s) for o, n, s in zip(offs, sizes, shapes)]
        self._choose_mode(bufs[0], grp)
        return bufs

    def _choose_mode(self, buf, grp):
        """Plan-time choice of the exchange kernel's form: multicast (in-switch reduction, (G+1)/G of the buffer per link
        direction) or peer loads / stores ((G-1)/G, but G loads per element), with 64 or 128 CTAs.  Every candidate is run on
        the largest exchange buffer and timed with CUDA events, max over ranks; every rank takes the same decision.
        LGB_EXCHANGE_MODE=multicast|peer and LGB_EXCHANGE_BLOCKS=n pin it."""
        import os
        want = os.environ.get("LGB_EXCHANGE_MODE", "auto")
        blocks = int(os.environ.get("LGB_EXCHANGE_BLOCKS", "0"))
        self.mode_report = {"multicast_available": self.multicast}
        modes = [m for m in ("multicast", "peer") if (m == "peer" or self.multicast) and want in ("auto", m)] or ["peer"]
        cands = [(m, b) for m in modes for b in ((blocks,) if blocks else (64, 128))]

        def flags_of(mode, b):
            return ((b & 0xFF) << 8) | (2 if mode == "peer" else 0)
        if len(cands) > 1:
            comm = self.comm[0]
            times = {}
            for mode, b in cands:
                self.flags = flags_of(mode, b)
                with torch.cuda.stream(comm):
                    for _ in range(2):
                        self._exchange(buf, 0, comm)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(comm)
                    for _ in range(5):
                        self._exchange(buf, 0, comm)
                    e1.record(comm)
                e1.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) / 5], device=self.device, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=grp)
                times[f"{mode}/{b}"] = float(t)
            best = min(times, key=times.get)
            cands = [c for c in cands if f"{c[0]}/{c[1]}" == best]
            self.mode_report.update(ms=times, chosen=best, bytes=buf.numel() * 4)
            buf.zero_()
            torch.cuda.synchronize(self.device)
            dist.barrier(group=grp)
        self.mode, self.blocks = cands[0]
        self.flags = flags_of(*cands[0])

    def _exchange(self, t, channel, comm):
        off = t.data_ptr() - self._arena.data_ptr() if self._arena is not None else -1
        if off < 0 or off + t.numel() * 4 > self._arena.numel() * 4 or t.numel() % 4 != 0 or not t.is_contiguous():
            raise RuntimeError("SymmOps: the exchanged tensor must be a contiguous buffer from alloc_exchange with numel % 4 == 0")
        with torch.cuda.device(self.device):
            check(_lib.load().lgb_exchange_allreduce_f32(C.byref(self._x), off, t.numel(), channel, int(getattr(self, "flags", 0)), comm.cuda_stream),
                  "exchange_allreduce")
        _lib.count_launch()


def make_ops(device, group=None, exchange: str = "auto"):
    """'symm' (own kernel over symmetric memory), 'nccl', or 'auto' = symm when every rank can set it up, else nccl."""
    if exchange == "nccl":
        return CudaOps(device, group)
    if exchange == "symm":
        return SymmOps(device, group)
    ok = 1.0
    try:
        ops = SymmOps(device, group)
    except Exception:
        ops, ok = None, 0.0
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        t = torch.tensor([ok], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        ok = float(t)
    return ops if ok > 0.5 else CudaOps(device, group)


# --------------------------------------------------------------------------------------------
# the sharded engine
# --------------------------------------------------------------------------------------------
class ShardedLightGCN:
    """LightGCN fwd + BPR + bwd over a user-sharded graph.  Every rank passes the same (users, items) COO.

    After ``fused_step``: ``E_f_users`` / ``E_f_items`` (final embeddings: owned user rows, replicated item rows),
    ``grad_users`` / ``grad_items`` (dLoss/dE0; the item rows are bit-identical on every rank; ``grad_items`` lives in an
    exchange buffer that the NEXT step's forward overwrites -- consume it, e.g. with ``adam_step``, before that)."""

    def __init__(self, num_users: int, num_items: int, embedding_dim: int, num_iterations: int,
                 users: torch.Tensor, items: torch.Tensor, device, group=None, ops=None,
                 rank: Optional[int] = None, world: Optional[int] = None, init_tables=None,
                 schedule: str = "chains", exchange: str = "auto", max_batch: int = 4096):
        self.U, self.I, self.d, self.K = int(num_users), int(num_items), int(embedding_dim), int(num_iterations)
        self.device = torch.device(device)
        if schedule not in ("chains", "layer"):
            raise ValueError(f"schedule={schedule!r} (chains: the two layer chains on two streams; layer: one stream)")
        if exchange not in ("auto", "nccl", "symm"):
            raise ValueError(f"exchange={exchange!r}")
        self.schedule = schedule
        inited = dist.is_available() and dist.is_initialized()
        self.rank = rank if rank is not None else (dist.get_rank(group) if inited else 0)
        self.world = world if world is not None else (dist.get_world_size(group) if inited else 1)
        self.ops = ops if ops is not None else make_ops(self.device, group, exchange)
        users, items = users.to(self.device), items.to(self.device)

        udeg = torch.bincount(users, minlength=self.U)
        self.bounds = balanced_user_bounds(udeg, self.world)
        self.lo, self.hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.Ug = Ug = self.hi - self.lo
        row, col, lu, li = local_block(users, items, self.lo, self.hi)
        self.local_edges = int(lu.numel())
        # global degrees: users are wholly local; item degrees are summed over ranks
        ideg = torch.bincount(li, minlength=self.I).to(torch.float32)
        self.ops.all_reduce(ideg)
        deg = torch.cat([udeg[self.lo:self.hi].to(torch.float32), ideg])
        dinv = torch.where(deg > 0, 1.0 / torch.sqrt(deg), torch.zeros_like(deg))
        self.n = n = Ug + self.I
        self.g_users, self.g_items = self.ops.build_views(row, col, Ug, self.I, dinv)
        del row, col, lu, li

        # parameters: one local table (users then items); the item block is identical on every rank
        f32 = dict(dtype=torch.float32, device=self.device)
        I, d, K = self.I, self.d, self.K
        self.table = torch.empty(n, d, **f32)
        if init_tables is not None:
            Wu, Wi = init_tables
            self.table[:Ug].copy_(Wu[self.lo:self.hi])
            self.table[Ug:].copy_(Wi)
        else:
            gen = torch.Generator(device=self.device).manual_seed(1000 + self.rank)
            self.table[:Ug].normal_(0, 0.1, generator=gen)
            gen_i = torch.Generator(device=self.device).manual_seed(999)
            self.table[Ug:].normal_(0, 0.1, generator=gen_i)
        self.max_batch = int(max_batch)
        self.rowsparse_backward = os.environ.get("LGB_ROWSPARSE_BACKWARD", "1") != "0"   # first backward layer: lgb_spmm_rowsparse ...
        self.rowsparse_ratio = 8                                                          # ... while the table has >= 8 rows per batch row
        # layer outputs: user rows in plain HBM, item rows in exchange buffers (+ the 2*B*d batch-row buffer of the BPR section)
        self._yu = [torch.empty(Ug, d, **f32) for _ in range(K)]
        bufs = self.ops.alloc_exchange([(I, d)] * max(K, 1) + [(2, self.max_batch, d)])
        self._yi, self._stage_full = bufs[:-1], bufs[-1]
        self.E_f_users, self.E_f_items = torch.empty(Ug, d, **f32), torch.empty(I, d, **f32)
        self.grad_users = torch.empty(Ug, d, **f32)
        self.grad_items = self._yi[K - 1] if K > 0 else torch.empty(I, d, **f32)
        self._ru, self._ri = torch.empty(Ug, d, **f32), torch.empty(I, d, **f32)
        self._dstage = torch.empty(self.max_batch, d, **f32)
        self._iota = torch.arange(self.max_batch, dtype=torch.int64, device=self.device)
        self._loss = torch.zeros((), **f32)
        self.loss = self._loss

    # ---- views kept for callers of the round-1 layout (tests, tools): concatenations, NOT the live buffers --------------
    @property
    def E_f(self):
        return torch.cat([self.E_f_users, self.E_f_items])

    @property
    def grad(self):
        return torch.cat([self.grad_users, self.grad_items])

    @property
    def users_weight(self):
        return self.table[: self.Ug]

    @property
    def items_weight(self):
        return self.table[self.Ug:]

    def graphs(self):
        return [self.g_users, self.g_items]

    def autotune(self):
        """Per-rank plan-time choice of the SpMM kernel variant / slice size for each view (DeviceCSR.autotune); ranks may
        choose differently -- the item rows are summed by the exchange, so the replicated blocks stay identical."""
        chunks = (2048, DEFAULT_CHUNK, 512, 256)      # a rank's launch is bounded below by one slice's serial chain: try shorter slices
        # (and longer ones for the CTA-wide-slice variants, which quarter the partial rows of stage 2)
        out = {}
        for name, g in (("users", self.g_users), ("items", self.g_items)):
            # the item rows gather from the local user shard (H&M shape: 351 MB / world, beyond the L2 up to 4 ranks): also try
            # their long-row slices in column-sweep order (csr.py use_sweep_order); the user rows gather from the L2-resident item block
            # degree-bucket row order: 18 % on the user rows at N = 1 (profiles/README.md r2u), tried on both views
            g.autotune(self.d, fused_epilogue=False, chunks=chunks, sweeps=(False, True) if name == "items" else (False,),
                       degree_orders=(False, True))
            out[name] = dict(g.autotune_report.get("chosen", {"variant": g.variant}), ms=g.autotune_report["ms"],
                             rejected=g.autotune_report["rejected"])
        return out

    # ---- optimizer step on the local shard (run_pipeline_lightgcn.py:103,159 -- optim.Adam over both tables) ------------
    def adam_step(self, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8) -> None:
        """One fused Adam step (lgb_adam_step, same arithmetic as torch.optim.Adam) on the local table with the gradients the
        last ``fused_step`` left behind: the owned user rows are updated by their owner only; the item rows carry identical
        gradients on every rank, so the replicated item block stays bit-identical without a broadcast."""
        if not hasattr(self, "_adam"):
            self._adam = dict(step=0, m=torch.zeros_like(self.table), v=torch.zeros_like(self.table))
        st, Ug = self._adam, self.Ug
        st["step"] += 1
        hp = (float(lr), float(betas[0]), float(betas[1]), float(eps), st["step"])
        if Ug:
            self.ops.adam(self.table[:Ug], self.grad_users, st["m"][:Ug], st["v"][:Ug], *hp)
        self.ops.adam(self.table[Ug:], self.grad_items, st["m"][Ug:], st["v"][Ug:], *hp)

    def autotune_step(self, user_indices, pos_item_indices, neg_item_indices, lambda_val: float, reps: int = 5,
                      candidates=("chains", "layer"), timer=None) -> dict:
        """Plan-time choice of the schedule on THIS machine: every candidate must reproduce the first candidate's loss and
        gradients on the given batch on EVERY rank (agreement by a MIN all-reduce) and is then timed (barrier, CUDA events,
        max over ranks); the fastest is kept.  Every rank runs the same candidates in the same order."""
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
        keep = self.schedule
        report, ref, best, best_ms = {"ms": {}, "rejected": {}}, None, keep, float("inf")
        for schedule in candidates:
            self.schedule = schedule
            loss = self.fused_step(user_indices, pos_item_indices, neg_item_indices, lambda_val).clone()
            grad = self.grad
            if ref is None:
                ref = (loss, grad)
            else:
                same = bool(torch.allclose(loss, ref[0], rtol=1e-5, atol=1e-7)) and \
                    bool(torch.allclose(grad, ref[1], rtol=1e-4, atol=1e-6 * float(ref[1].abs().max()) + 1e-12))
                if agree(1.0 if same else 0.0, dist.ReduceOp.MIN) < 0.5:
                    report["rejected"][schedule] = "loss / gradients differ from the first candidate on at least one rank"
                    continue
            ms = agree(timer(lambda: self.fused_step(user_indices, pos_item_indices, neg_item_indices, lambda_val)), dist.ReduceOp.MAX)
            report["ms"][schedule] = ms
            if ms < best_ms:
                best, best_ms = schedule, ms
        self.schedule = best
        report["chosen"] = {"schedule": best}
        return report

    # ---- K propagation layers as two independent chains ---------------------------------------------------------------
    def _propagate(self, xu, xi, resid_u=None, resid_i=None, last_u=None, rows_u=None, rows_i=None):
        """(xu, xi) -> K layers of  yu = G_users xi (+resid_u),  yi = exchange(G_items xu (+resid_i on rank 0)).
        Returns ([yu_0..yu_{K-1}], [yi_0..yi_{K-1}]); on return everything is ordered on the current stream.
        last_u: buffer that receives yu_{K-1} instead of the internal one.  rows_u / rows_i: bitmaps of the rows of the INPUT
        xu / xi that may be non-zero (the backward's batch rows): layer 0 gathers from them as lgb_spmm_rowsparse."""
        K, ops = self.K, self.ops
        two = self.schedule == "chains"
        ctx = ops.fork(two)
        yus, yis, pending = [], [], None
        ri = resid_i if self.rank == 0 else None          # delivered exactly once through the sum over ranks
        for k in range(K):
            ci, cu = (1, 0) if k % 2 == 0 else (0, 1)      # chain of I(k) / of U(k)
            yi = self._yi[k]
            yu = last_u if (k == K - 1 and last_u is not None) else self._yu[k]
            with ctx[ci]:
                ops.spmm(self.g_items, xu, yi, resid=ri, x_rows=rows_u if k == 0 else None,
                         resid_rows=rows_i if k == 0 else None)                 # partial item rows <- owned users
                h = ops.exchange_async(yi, channel=ci)                           # ... summed over ranks
            with ctx[cu]:
                if pending is not None:
                    pending.wait()                                               # xi = yi_{k-1} must be complete (same chain)
                ops.spmm(self.g_users, xi, yu, resid=resid_u, x_rows=rows_i if k == 0 else None,
                         resid_rows=rows_u if k == 0 else None)                 # owned user rows <- all items
            yus.append(yu); yis.append(yi)
            xu, xi, pending = yu, yi, h
        ops.join(two)
        if pending is not None:
            pending.wait()
        return yus, yis

    def forward(self):
        """E_f = mean_k A^k E0 on the local rows (item rows replicated) -> (E_f_users, E_f_items)."""
        K, Ug, ops = self.K, self.Ug, self.ops
        Wu, Wi = self.table[:Ug], self.table[Ug:]
        if K == 0:
            self.E_f_users.copy_(Wu); self.E_f_items.copy_(Wi)
            return self.E_f_users, self.E_f_items
        yus, yis = self._propagate(Wu, Wi)
        if Ug:
            ops.mean_rows([Wu] + yus, float(K + 1), self.E_f_users)
        ops.mean_rows([Wi] + yis, float(K + 1), self.E_f_items)
        return self.E_f_users, self.E_f_items

    def backward(self, ru: torch.Tensor, ri: torch.Tensor, rows_u=None, rows_i=None):
        """grad = sum_k (A^T)^k r with r = dE_f/(K+1) (Horner: g <- A g + r; the local block is symmetric).  ri must be the
        complete item-row residual (identical on every rank).  rows_u / rows_i: optional bitmaps (ops.rows_bitmap) of the rows
        of ru / ri that may be non-zero."""
        if self.K == 0:
            # no propagation: the item-row gradient is the BPR scatter itself.  Every rank computed it, but with atomics, whose
            # order (hence the last bit of a row that received three or more contributions) is not the same everywhere: all
            # ranks adopt rank 0's bits through the exchange (x + 0 + ... + 0), like the K >= 1 path does for its residual
            self.grad_users.copy_(ru)
            yi = self._yi[0]
            if self.rank == 0:
                yi.copy_(ri)
            else:
                self.ops.zero(yi)
            self.ops.exchange_async(yi, channel=0).wait()
            self.grad_items.copy_(yi)
            return self.grad_users, self.grad_items
        self._propagate(ru, ri, resid_u=ru, resid_i=ri, last_u=self.grad_users, rows_u=rows_u, rows_i=rows_i)
        return self.grad_users, self.grad_items

    @torch.no_grad()
    def fused_step(self, user_indices: torch.Tensor, pos_item_indices: torch.Tensor, neg_item_indices: torch.Tensor,
                   lambda_val: float) -> torch.Tensor:
        """Global batch in (the same B triples on every rank), global loss out (0-dim tensor, bit-identical on every rank);
        gradients in ``grad_users`` (owned rows) / ``grad_items`` (replicated rows, identical on every rank).  Static
        shapes, no host synchronisation: CUDA-graph capturable (``capture``)."""
        ops, Ug, K, d = self.ops, self.Ug, self.K, self.d
        u, p, n = (_lib.i64c(t.to(self.device)) for t in (user_indices, pos_item_indices, neg_item_indices))
        B = u.numel()
        if B > self.max_batch:
            raise RuntimeError(f"batch of {B} triples > max_batch={self.max_batch} (the symmetric batch-row buffer is sized at construction)")
        Wu, Wi = self.table[:Ug], self.table[Ug:]
        Efu, Efi = self.forward()
        # the batch's user rows, gathered by their owners: stage[0] = E_f rows, stage[1] = layer-0 rows
        stage = self._stage_full.view(-1)[: 2 * B * d].view(2, B, d)
        ops.gather_owned(Efu, u, self.lo, self.hi, stage[0])
        ops.gather_owned(Wu, u, self.lo, self.hi, stage[1])
        h = ops.exchange_async(stage, channel=2)
        ru, ri, dst = self._ru, self._ri, self._dstage[:B]
        ops.zero(ru); ops.zero(ri); ops.zero(dst)
        h.wait()
        # every rank evaluates ALL triples: loss and item-row gradients complete and identical everywhere
        loss = self._loss
        ops.bpr(stage[0], stage[1], Efi, Wi, self._iota[:B], p, n, lambda_val, B, gscale=1.0 / (K + 1), loss=loss, duf=dst, dpf=ri)
        ops.scatter_add_owned(dst, u, self.lo, self.hi, ru)
        # ru / ri are non-zero on the batch rows only: the first backward layer skips every entry that multiplies a zero row
        rows_u = rows_i = None
        if self.rowsparse_backward and d % 4 == 0 and d <= 64:
            if self.rowsparse_ratio * B <= Ug:
                rows_u = ops.rows_bitmap(Ug, ((u, -self.lo),), "users")
            if self.rowsparse_ratio * 2 * B <= self.I:
                rows_i = ops.rows_bitmap(self.I, ((p, 0), (n, 0)), "items")
        Gu, Gi = self.backward(ru, ri, rows_u=rows_u, rows_i=rows_i)
        # + 2*lambda*E0 on the batch rows: owned users locally, item rows redundantly (identical) on every rank
        ops.bpr(Efu, Wu, Efi, Wi, u, p, n, lambda_val, B, user_lo=self.lo, user_hi=self.hi, user_rows_only=True, du0=Gu, dp0=Gi)
        self.loss = loss
        return loss

    # ---- CUDA-graph replay of the whole step (kernels on three streams + the exchange kernels) --------------------------
    def capture(self, batch_size: int, lambda_val: float):
        """Capture the step for a fixed batch size; returns step(u, p, n) -> loss that replays the graph."""
        dev = self.device
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
