"""CPU tests (gloo, world_size 2 and 3) of the multi-GPU host logic in laplace_gnn_recommendation_b200/dist.py:
nnz-balanced user partition, local symmetric blocks with GLOBAL normalisation, the per-layer item-block
all-reduce, sharded BPR with partial item gradients.  The kernels are replaced by an oracle-backed ops object
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

    def build_graph(self, row, col, n, dinv):
        rowptr, c, _ = lo.csr_from_coo(row, col, n, n)
        r = lo.rows_from_rowptr(rowptr)
        return dict(rowptr=rowptr, col=c, val=(1.0 * dinv[r]) * dinv[c], n=n)

    def row_view(self, g, lo_, hi):
        s, e = int(g["rowptr"][lo_]), int(g["rowptr"][hi])
        return dict(rowptr=g["rowptr"][lo_:hi + 1] - s, col=g["col"][s:e], val=g["val"][s:e], n=hi - lo_)

    def spmm(self, g, X, Y=None, resid=None, acc_in=None, acc_out=None, acc_div=1.0):
        y = lo.spmm(g["rowptr"], g["col"], g["val"], X)
        if resid is not None:
            y = y + resid
        if acc_out is not None:
            acc_out.copy_(((acc_in if acc_in is not None else 0) + y) / acc_div)
        if Y is not None:
            Y.copy_(y)

    def spmm_split(self, g, X, split_row, y_tail, Y=None, resid=None, acc_in=None, acc_out=None, acc_div=1.0):
        y = lo.spmm(g["rowptr"], g["col"], g["val"], X)
        y_tail.copy_(y[split_row:])
        yu = y[:split_row] if resid is None else y[:split_row] + resid[:split_row]
        if acc_out is not None:
            acc_out[:split_row].copy_(((acc_in[:split_row] if acc_in is not None else 0) + yu) / acc_div)
        if Y is not None:
            Y[:split_row].copy_(yu)

    def accumulate(self, y, acc, resid, div, out):
        v = y if resid is None else y + resid
        if acc is not None:
            v = acc + v
        out.copy_(v / div)

    def zero(self, t):
        t.zero_()

    def bpr(self, Ef, E0, Ug, u, p, n, lam, B_norm, user_lo=0, user_hi=0, loss=None, dEf=None, dE0_users=None,
            dE0_items=None, gscale=1.0):
        B = B_norm
        if user_hi > 0:
            mine = (u >= user_lo) & (u < user_hi)
            u, p, n = u[mine] - user_lo, p[mine], n[mine]
        uf, pf, nf = Ef[u], Ef[Ug + p], Ef[Ug + n]
        u0, p0, n0 = E0[u], E0[Ug + p], E0[Ug + n]
        x = (uf * pf).sum(-1) - (uf * nf).sum(-1)
        if loss is not None:
            loss.copy_(-F.softplus(x).sum() / B + lam * ((u0 ** 2).sum() + (p0 ** 2).sum() + (n0 ** 2).sum()))
        c = (-torch.sigmoid(x) / B * gscale).unsqueeze(1)
        if dEf is not None:
            dEf.index_add_(0, u, c * (pf - nf))
            dEf.index_add_(0, Ug + p, c * uf)
            dEf.index_add_(0, Ug + n, -c * uf)
        if dE0_users is not None:
            dE0_users.index_add_(0, u, 2 * lam * u0)
        if dE0_items is not None:
            dE0_items.index_add_(0, p, 2 * lam * p0)
            dE0_items.index_add_(0, n, 2 * lam * n0)

    def adam(self, p, g, m, v, lr, beta1, beta2, eps, step):
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
        p.addcdiv_(m, (v.sqrt() / (bc2 ** 0.5)).add_(eps), value=-lr / bc1)

    class _Done:
        def wait(self):
            pass

    def all_reduce_async(self, t):
        self.all_reduce(t)
        return self._Done()

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
    """dist.CudaOps itself -- row views, lgb_gcn_values, lgb_spmm(_split), lgb_accumulate, lgb_bpr with B_norm and the
    owned-user filter -- with the kernels served by the CPU emulation of the library (tests/emu/) and the collectives by
    gloo.  Only the two CUDA-stream specifics of the class are replaced."""
    from laplace_gnn_recommendation_b200.dist import CudaOps, _Done

    class _Lazy:
        """Handle of an all-reduce that only happens when somebody waits for it: the latest moment a real asynchronous
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
            self.device, self.group, self.comm, self._bpr_ws = device, group, None, None

        def all_reduce_async(self, t):
            if lazy:
                return _Lazy(self, t)
            self.all_reduce(t)
            return _Done()

    return EmuOps(torch.device("cpu"))


def _worker(rank, world, port, cases, out_dir):
    """One rank of a gloo group: runs every case (K, schedule, static_batch, ops_kind, d) in turn -- one process start-up
    per world size instead of one per case."""
    import contextlib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        for ci, (K, schedule, static_batch, ops_kind, d) in enumerate(cases):
            pb = make_problem(K=K, d=d)
            if ops_kind == "emu":
                from tests.emu.harness import emulated
                ctx = emulated()
            else:
                ctx = contextlib.nullcontext()
            with ctx:
                ops = make_emu_ops() if ops_kind == "emu" else CpuOracleOps()
                eng = ShardedLightGCN(pb["U"], pb["I"], pb["d"], pb["K"], pb["users"], pb["items"], "cpu", ops=ops,
                                      init_tables=(pb["Wu"], pb["Wi"]), schedule=schedule, static_batch=static_batch)
                loss = eng.fused_step(pb["u"], pb["p"], pb["n"], pb["lam"])
                tuned = None
                if ops_kind == "emu" and ci == len(cases) - 1:      # plan-time choice of the step form, once per world size
                    import time
                    tuned = eng.autotune_step(pb["u"], pb["p"], pb["n"], pb["lam"], candidates=((None, False), (None, True), ("pipelined", True)),
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


# (K, schedule, static_batch, ops_kind, d).  "oracle": the kernels are the CPU oracle (host logic only); "emu": dist.CudaOps
# drives the REAL kernel sources under the CPU emulator (tests/emu/) -- covers lgb_spmm_split, the BPR owned-user filter /
# B_norm and the row views under every schedule without a GPU, with lazy all-reduce handles (see make_emu_ops).
CASES = {
    2: [(3, "layer", False, "oracle", 16), (1, "layer", True, "oracle", 16), (3, "pipelined", True, "oracle", 16),
        (3, "merged", True, "oracle", 16), (1, "merged", False, "oracle", 16), (1, "pipelined", True, "oracle", 16),
        (3, "layer", False, "emu", 64), (3, "merged", True, "emu", 64), (2, "merged", False, "emu", 32),
        (3, "pipelined", False, "emu", 64), (0, "layer", True, "emu", 64)],
    3: [(2, "layer", False, "oracle", 16), (2, "pipelined", False, "oracle", 16), (2, "merged", False, "oracle", 16),
        (2, "pipelined", True, "emu", 64), (1, "layer", True, "emu", 128), (2, "merged", True, "emu", 32)],
}


@pytest.mark.parametrize("world", sorted(CASES))
def test_sharded_step_equals_single_process_oracle(tmp_path, world):
    """Every rank's shard of one fused step (loss, E_f, gradients) equals the single-process oracle, for every schedule /
    static-batch combination, with oracle-backed ops and with the emulated kernels."""
    from tests.emu import build_emu
    build_emu.build()          # once, in the parent: the spawned ranks only load it
    cases = CASES[world]
    mp.spawn(_worker, args=(world, _free_port(), cases, str(tmp_path)), nprocs=world, join=True)
    for ci, (K, schedule, static_batch, ops_kind, d) in enumerate(cases):
        what = f"world={world} K={K} schedule={schedule} static={static_batch} ops={ops_kind} d={d}"
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
            assert len(t0["ms"]) == 3 and not t0["rejected"] and all(o["tuned"] == t0 for o in outs), t0
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
