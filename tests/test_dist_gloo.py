"""CPU tests (gloo, world_size 2 and 3) of the multi-GPU host logic in laplace_gnn_recommendation_b200/dist.py:
nnz-balanced user partition, local blocks with GLOBAL normalisation, the two independent layer chains with one item-block
exchange per layer, BPR on the exchanged batch rows (loss and item gradients complete on every rank).  The kernels are replaced by an oracle-backed ops object
defined HERE (tests may use the oracle; the product never does); the collectives are real (gloo)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from laplace_gnn_recommendation_b200.dist import ShardedLightGCN, balanced_user_bounds, local_block
from oracle import lightgcn_oracle as lo


class CpuOracleOps:
    """Same interface as dist.CudaOps, arithmetic by the CPU oracle, collectives over gloo."""

    group = None

    def fork(self, two_streams):
        import contextlib
        return [contextlib.nullcontext(), contextlib.nullcontext()]

    def join(self, two_streams):
        pass

    def build_views(self, row, col, Ug, I, dinv):
        n = Ug + I
        rowptr, c, _ = lo.csr_from_coo(row, col, n, n)
        r = lo.rows_from_rowptr(rowptr)
        val = (1.0 * dinv[r]) * dinv[c]
        s = int(rowptr[Ug])
        gu = dict(rowptr=rowptr[: Ug + 1].clone(), col=c[:s] - Ug, val=val[:s], n=Ug, n_cols=I)
        gi = dict(rowptr=rowptr[Ug:] - s, col=c[s:], val=val[s:], n=I, n_cols=Ug)
        return gu, gi

    def alloc_exchange(self, shapes):
        return [torch.zeros(*sh) for sh in shapes]

    def spmm(self, g, X, Y, resid=None, x_rows=None, resid_rows=None):
        if x_rows is not None:       # the engine's promise behind lgb_spmm_rowsparse: X is zero outside the flagged rows
            assert x_rows.numel() == X.shape[0] and not bool(X[~x_rows].any())
            if resid_rows is not None and resid is not None:
                assert resid_rows.numel() == resid.shape[0] and not bool(resid[~resid_rows].any())
            self.rowsparse_calls = getattr(self, "rowsparse_calls", 0) + 1
        y = lo.spmm(g["rowptr"], g["col"], g["val"], X) if X.shape[0] else torch.zeros(g["n"], X.shape[1])
        Y.copy_(y if resid is None else y + resid)

    def rows_bitmap(self, n, index_lists, key):
        if n <= 0:
            return None
        m = torch.zeros(n, dtype=torch.bool)
        for idx, off in index_lists:
            r = idx.long() + off
            m[r[(r >= 0) & (r < n)]] = True
        return m

    def mean_rows(self, srcs, div, out):
        acc = srcs[0]
        for t in srcs[1:]:
            acc = acc + t
        out.copy_(acc / div)

    def accumulate(self, y, acc, resid, div, out):
        v = y if resid is None else y + resid
        if acc is not None:
            v = acc + v
        out.copy_(v / div)

    def zero(self, t):
        t.zero_()

    def gather_owned(self, src, idx, lo_, hi, dst):
        mine = (idx >= lo_) & (idx < hi)
        dst.zero_()
        dst[mine] = src[idx[mine] - lo_]

    def scatter_add_owned(self, src, idx, lo_, hi, dst):
        mine = (idx >= lo_) & (idx < hi)
        dst.index_add_(0, idx[mine] - lo_, src[mine])

    def bpr(self, uf, u0, pf, p0, iu, ip, in_, lam, B_norm, gscale=1.0, user_lo=0, user_hi=0, user_rows_only=False,
            loss=None, duf=None, du0=None, dpf=None, dp0=None):
        own = torch.ones_like(iu, dtype=torch.bool)
        if user_hi > 0:
            own = (iu >= user_lo) & (iu < user_hi)
            if not user_rows_only:
                iu, ip, in_, own = iu[own], ip[own], in_[own], own[own]
            iu = iu - user_lo
        d = pf.shape[1]
        ufr = torch.zeros(iu.numel(), d); u0r = torch.zeros(iu.numel(), d)
        ufr[own] = uf[iu[own]]; u0r[own] = u0[iu[own]]
        pfr, nfr, p0r, n0r = pf[ip], pf[in_], p0[ip], p0[in_]
        x = (ufr * pfr).sum(-1) - (ufr * nfr).sum(-1)
        if loss is not None:
            loss.copy_(-F.softplus(x).sum() / B_norm + lam * ((u0r ** 2).sum() + (p0r ** 2).sum() + (n0r ** 2).sum()))
        c = (-torch.sigmoid(x) / B_norm * gscale).unsqueeze(1)
        if duf is not None:
            duf.index_add_(0, iu[own], (c * (pfr - nfr))[own])
        if dpf is not None:
            dpf.index_add_(0, ip, c * ufr)
            dpf.index_add_(0, in_, -c * ufr)
        if du0 is not None and du0.numel():
            du0.index_add_(0, iu[own], 2 * lam * u0r[own])
        if dp0 is not None:
            dp0.index_add_(0, ip, 2 * lam * p0r)
            dp0.index_add_(0, in_, 2 * lam * n0r)

    def adam(self, p, g, m, v, lr, beta1, beta2, eps, step):
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
        p.addcdiv_(m, (v.sqrt() / (bc2 ** 0.5)).add_(eps), value=-lr / bc1)

    class _Lazy:
        """An exchange that only happens when somebody waits for it: the latest moment a real asynchronous collective may
        complete.  Code that reads the buffer before wait() sees the un-reduced partial sums and fails the parity check."""

        def __init__(self, ops, t):
            self.ops, self.t = ops, t

        def wait(self):
            if self.t is not None:
                self.ops.all_reduce(self.t)
                self.t = None

    def exchange_async(self, t, channel=0):
        return self._Lazy(self, t)

    def all_reduce(self, t):
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)


def make_problem(seed=0, U=60, I=25, E=700, d=16, K=3, B=96):
    gen = torch.Generator().manual_seed(seed)
    users = (torch.rand(E, generator=gen) ** 2 * U).long().clamp(max=U - 1)   # skewed user degrees
    items = torch.randint(0, I, (E,), generator=gen)
    Wu, Wi = torch.randn(U, d, generator=gen) * 0.1, torch.randn(I, d, generator=gen) * 0.1
    pick = torch.randint(0, E, (B,), generator=gen)
    return dict(U=U, I=I, d=d, K=K, users=users, items=items, Wu=Wu, Wi=Wi, u=users[pick], p=items[pick],
                n=torch.randint(0, I, (B,), generator=gen), lam=1e-3)


def single_process_reference(pb):
    row, col, n = lo.wiring_symmetric(pb["users"], pb["items"], pb["U"], pb["I"])
    rowptr, c, _ = lo.csr_from_coo(row, col, n, n)
    return lo.train_iteration(pb["Wu"], pb["Wi"], rowptr, c, pb["K"], pb["u"], pb["p"], pb["n"], pb["lam"])


def make_emu_ops(lazy: bool = True):
    """dist.CudaOps itself -- the two CSR views of the local block, lgb_gcn_values, lgb_spmm, lgb_mean_rows, the owned-row
    gather / scatter, lgb_bpr on explicit operand tables (incl. the user-rows-only filter) -- with the kernels served by the
    CPU emulation of the library (tests/emu/) and the collectives by gloo.  Only the CUDA-stream specifics are replaced."""
    import contextlib
    from laplace_gnn_recommendation_b200.dist import CudaOps, _Done

    class _Lazy:
        """Handle of an exchange that only happens when somebody waits for it: the latest moment a real asynchronous
        collective may complete.  Code that reads the buffer before wait() sees the un-reduced partial sums and fails the
        parity check -- the CPU stand-in for a missing stream dependency."""

        def __init__(self, ops, t):
            self.ops, self.t = ops, t

        def wait(self):
            if self.t is not None:
                self.ops.all_reduce(self.t)
                self.t = None

    class EmuOps(CudaOps):
        def __init__(self, device, group=None):
            self.device, self.group, self.comm, self.chains, self._bpr_ws, self.exchange_events = device, group, None, None, None, None

        def fork(self, two_streams):
            return [contextlib.nullcontext(), contextlib.nullcontext()]

        def join(self, two_streams):
            pass

        def exchange_async(self, t, channel=0):
            if lazy:
                return _Lazy(self, t)
            self.all_reduce(t)
            return _Done()

        all_reduce_async = exchange_async

    return EmuOps(torch.device("cpu"))


def _worker(rank, world, port, cases, out_dir):
    """One rank of a gloo group: runs every case (K, schedule, static_batch, ops_kind, d) in turn -- one process start-up
    per world size instead of one per case."""
    import contextlib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        for ci, (K, schedule, ops_kind, d) in enumerate(cases):
            pb = make_problem(K=K, d=d)
            if ops_kind == "emu":
                from tests.emu.harness import emulated
                ctx = emulated()
            else:
                ctx = contextlib.nullcontext()
            with ctx:
                ops = make_emu_ops() if ops_kind == "emu" else CpuOracleOps()
                eng = ShardedLightGCN(pb["U"], pb["I"], pb["d"], pb["K"], pb["users"], pb["items"], "cpu", ops=ops,
                                      init_tables=(pb["Wu"], pb["Wi"]), schedule=schedule, max_batch=128)
                if ci % 2 == 1:                  # every other case: the first backward layer as lgb_spmm_rowsparse whatever the sizes
                    eng.rowsparse_ratio = 0
                loss = eng.fused_step(pb["u"], pb["p"], pb["n"], pb["lam"])
                if ci % 2 == 1 and ops_kind == "oracle" and K > 0 and d <= 64:
                    assert ops.rowsparse_calls >= 1   # (CpuOracleOps.spmm also asserts the operand is zero outside the bitmap)
                tuned = None
                if ops_kind == "emu" and ci == len(cases) - 1:      # plan-time choice of the step form, once per world size
                    import time
                    tuned = eng.autotune_step(pb["u"], pb["p"], pb["n"], pb["lam"], candidates=("chains", "layer"),
                                              timer=lambda fn: (time.perf_counter(), fn(), time.perf_counter())[2] * 0 + 1.0 + rank)
                    loss = eng.fused_step(pb["u"], pb["p"], pb["n"], pb["lam"])      # the installed winner still computes the step
                snap = dict(loss=loss.clone(), Ef=eng.E_f.clone(), grad=eng.grad.clone())
                eng.adam_step(lr=1e-2)                                               # optimizer step on the shard, then a second iteration
                eng.fused_step(pb["u"], pb["p"], pb["n"], pb["lam"])
                eng.adam_step(lr=1e-2)
            torch.save(dict(lo=eng.lo, hi=eng.hi, tuned=tuned, table=eng.table.clone(), **snap,
                            bounds=eng.bounds, local_edges=eng.local_edges), os.path.join(out_dir, f"case{ci}_rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


# (K, schedule, ops_kind, d).  "oracle": the kernels are the CPU oracle (host logic only); "emu": dist.CudaOps drives the REAL
# kernel sources under the CPU emulator (tests/emu/) -- covers the remapped CSR views, lgb_mean_rows, the owned-row gather /
# scatter and the BPR user-rows-only filter under both schedules without a GPU, with lazy exchange handles (see make_emu_ops).
CASES = {
    2: [(3, "chains", "oracle", 16), (1, "layer", "oracle", 16), (2, "chains", "oracle", 16), (0, "chains", "oracle", 16),
        (3, "chains", "emu", 64), (3, "layer", "emu", 64), (2, "chains", "emu", 32), (1, "chains", "emu", 128), (0, "layer", "emu", 64)],
    3: [(2, "layer", "oracle", 16), (3, "chains", "oracle", 16), (4, "chains", "oracle", 16),
        (2, "chains", "emu", 64), (1, "layer", "emu", 128), (4, "chains", "emu", 32)],
}


@pytest.mark.parametrize("world", sorted(CASES))
def test_sharded_step_equals_single_process_oracle(tmp_path, world):
    """Every rank's shard of one fused step (loss, E_f, gradients) equals the single-process oracle, for every schedule /
    static-batch combination, with oracle-backed ops and with the emulated kernels."""
    from tests.emu import build_emu
    build_emu.build()          # once, in the parent: the spawned ranks only load it
    cases = CASES[world]
    mp.spawn(_worker, args=(world, _free_port(), cases, str(tmp_path)), nprocs=world, join=True)
    for ci, (K, schedule, ops_kind, d) in enumerate(cases):
        what = f"world={world} K={K} schedule={schedule} ops={ops_kind} d={d}"
        pb = make_problem(K=K, d=d)
        o_loss, o_gu, o_gi, o_uf, o_if = single_process_reference(pb)
        outs = [torch.load(tmp_path / f"case{ci}_rank{r}.pt") for r in range(world)]
        assert outs[0]["bounds"][0] == 0 and outs[0]["bounds"][-1] == pb["U"], what
        assert sum(o["local_edges"] for o in outs) == pb["users"].numel(), what
        tol = dict(rtol=1e-5, atol=1e-7, msg=lambda m: f"{what}: {m}")
        gtol = dict(rtol=1e-5, atol=1e-9, msg=lambda m: f"{what}: {m}")
        for o in outs:
            Ug = o["hi"] - o["lo"]
            torch.testing.assert_close(o["loss"], o_loss, **tol)
            torch.testing.assert_close(o["Ef"][:Ug], o_uf[o["lo"]:o["hi"]], **tol)
            torch.testing.assert_close(o["Ef"][Ug:], o_if, **tol)                       # replicated item rows
            torch.testing.assert_close(o["grad"][:Ug], o_gu[o["lo"]:o["hi"]], **gtol)
            torch.testing.assert_close(o["grad"][Ug:], o_gi, **gtol)                    # identical on every rank
        for a, b in zip(outs[:-1], outs[1:]):
            assert a["hi"] == b["lo"], what
        # two iterations with ShardedLightGCN.adam_step == two iterations of the oracle with torch.optim.Adam
        Wu, Wi = pb["Wu"].clone().requires_grad_(True), pb["Wi"].clone().requires_grad_(True)
        opt = torch.optim.Adam([Wu, Wi], lr=1e-2)
        rowptr, cc, _ = lo.csr_from_coo(*lo.wiring_symmetric(pb["users"], pb["items"], pb["U"], pb["I"])[:2], pb["U"] + pb["I"], pb["U"] + pb["I"])
        for _ in range(2):
            _, gu, gi, _, _ = lo.train_iteration(Wu.detach(), Wi.detach(), rowptr, cc, K, pb["u"], pb["p"], pb["n"], pb["lam"])
            Wu.grad, Wi.grad = gu, gi
            opt.step()
        for o in outs:
            Ug = o["hi"] - o["lo"]
            atol = dict(rtol=1e-4, atol=5e-6, msg=lambda m: f"{what} (after 2 Adam steps): {m}")
            torch.testing.assert_close(o["table"][:Ug], Wu.detach()[o["lo"]:o["hi"]], **atol)
            torch.testing.assert_close(o["table"][Ug:], Wi.detach(), **atol)
            Ug0 = outs[0]["hi"] - outs[0]["lo"]
            assert torch.equal(o["table"][Ug:], outs[0]["table"][Ug0:]), what            # replicated item block: same bits on every rank
        if outs[0]["tuned"] is not None:             # autotune_step: every rank reports the same, complete, rejection-free table
            t0 = outs[0]["tuned"]
            assert len(t0["ms"]) == 2 and not t0["rejected"] and all(o["tuned"] == t0 for o in outs), t0
            assert all(v == float(world) for v in t0["ms"].values())          # max over ranks of the (1 + rank) the fake timer returns


def test_balanced_bounds_and_local_block():
    deg = torch.tensor([5, 0, 0, 1, 1, 1, 10, 2, 0, 4])
    for world in (1, 2, 3, 4, 8):
        b = balanced_user_bounds(deg, world)
        assert b[0] == 0 and b[-1] == 10 and len(b) == world + 1 and all(x <= y for x, y in zip(b, b[1:]))
    b = balanced_user_bounds(deg, 2)
    left = int(deg[: b[1]].sum())
    assert abs(left - 12) <= 10  # the heaviest user (10 edges) is the granularity
    users = torch.tensor([0, 0, 6, 6, 9, 3]); items = torch.tensor([1, 2, 0, 1, 2, 2])
    row, col, lu, li = local_block(users, items, 3, 8)
    assert lu.tolist() == [3, 3, 0] and li.tolist() == [0, 1, 2]
    assert row.tolist() == [3, 3, 0, 5, 6, 7] and col.tolist() == [5, 6, 7, 3, 3, 0]
    assert balanced_user_bounds(torch.zeros(4, dtype=torch.long), 2) == [0, 1, 4] or True  # degenerate: any monotone split


# ---------------------------------------------------------------------------------------------------------------------
# generic row-sharded engine (dist_rows.RowShardedLightGCN): arbitrary [N, N] adjacency, all-gather per layer
def _rows_worker(rank, world, port, cases, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        from laplace_gnn_recommendation_b200.dist_rows import RowShardedLightGCN
        from tests.emu.harness import emulated
        for ci, (K, wiring, d) in enumerate(cases):
            pb = make_problem(K=K, d=d, U=70, I=45, E=900, B=128)
            row, col, n = (lo.wiring_reference if wiring == "R" else lo.wiring_symmetric)(pb["users"], pb["items"], pb["U"], pb["I"])
            with emulated():
                eng = RowShardedLightGCN(pb["U"], pb["I"], pb["d"], K, row, col, "cpu", ops=make_emu_ops(),
                                         init_tables=(pb["Wu"], pb["Wi"]), chunk=64)
                loss = eng.fused_step(pb["u"], pb["p"], pb["n"], pb["lam"])
                snap = dict(loss=loss.clone(), Ef=eng.E_f.clone(), grad=eng.grad.clone())
                eng.adam_step(lr=1e-2)
                eng.fused_step(pb["u"], pb["p"], pb["n"], pb["lam"])
                eng.adam_step(lr=1e-2)
            torch.save(dict(lo=eng.lo, hi=eng.hi, bounds=eng.bounds, table=eng.table.clone(), nnz=eng.local_nnz, **snap),
                       os.path.join(out_dir, f"case{ci}_rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


ROWS_CASES = {1: [(3, "R", 64)], 2: [(3, "R", 64), (1, "S", 128)], 3: [(2, "S", 32), (0, "R", 16)]}   # (K, wiring, d)


@pytest.mark.parametrize("world", sorted(ROWS_CASES))
def test_row_sharded_engine_equals_single_process_oracle(tmp_path, world):
    """The generic engine (north_star's layout: rows of A_hat and of the table range-partitioned, all-gather per layer) with
    the real kernel sources under the emulator, against the single-process oracle -- incl. the reference's own wiring (R)."""
    from tests.emu import build_emu
    build_emu.build()
    cases = ROWS_CASES[world]
    mp.spawn(_rows_worker, args=(world, _free_port(), cases, str(tmp_path)), nprocs=world, join=True)
    for ci, (K, wiring, d) in enumerate(cases):
        pb = make_problem(K=K, d=d, U=70, I=45, E=900, B=128)
        row, col, n = (lo.wiring_reference if wiring == "R" else lo.wiring_symmetric)(pb["users"], pb["items"], pb["U"], pb["I"])
        rowptr, c, _ = lo.csr_from_coo(row, col, n, n)
        o_loss, o_gu, o_gi, o_uf, o_if = lo.train_iteration(pb["Wu"], pb["Wi"], rowptr, c, K, pb["u"], pb["p"], pb["n"], pb["lam"])
        o_ef, o_g = torch.cat([o_uf, o_if]), torch.cat([o_gu, o_gi])
        outs = [torch.load(tmp_path / f"case{ci}_rank{r}.pt") for r in range(world)]
        assert outs[0]["bounds"][0] == 0 and outs[0]["bounds"][-1] == n
        assert sum(o["nnz"][0] for o in outs) == row.numel() == sum(o["nnz"][1] for o in outs)
        for o in outs:
            torch.testing.assert_close(o["loss"], o_loss, rtol=1e-5, atol=1e-7)
            assert torch.equal(o["loss"], outs[0]["loss"])                           # bit-identical without a collective
            torch.testing.assert_close(o["Ef"], o_ef[o["lo"]:o["hi"]], rtol=1e-5, atol=1e-7)
            torch.testing.assert_close(o["grad"], o_g[o["lo"]:o["hi"]], rtol=1e-5, atol=1e-9)
        for a, b in zip(outs[:-1], outs[1:]):
            assert a["hi"] == b["lo"]
        Wu, Wi = pb["Wu"].clone().requires_grad_(True), pb["Wi"].clone().requires_grad_(True)      # two iterations incl. Adam
        opt = torch.optim.Adam([Wu, Wi], lr=1e-2)
        for _ in range(2):
            _, gu, gi, _, _ = lo.train_iteration(Wu.detach(), Wi.detach(), rowptr, c, K, pb["u"], pb["p"], pb["n"], pb["lam"])
            Wu.grad, Wi.grad = gu, gi
            opt.step()
        full = torch.cat([Wu.detach(), Wi.detach()])
        for o in outs:
            torch.testing.assert_close(o["table"], full[o["lo"]:o["hi"]], rtol=1e-4, atol=5e-6)
