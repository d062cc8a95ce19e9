"""Probe: time lgb_spmm on the H&M-shaped graph split into its two halves (user rows <- items, item rows <- users)
for a list of kernel variants.  python tools/spmm_probe.py [--degree powerlaw|uniform] [--variants 1,3,6] [--d 64]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _common  # noqa: E402
from _common import timeit  # noqa: E402
import laplace_gnn_recommendation_b200 as lg  # noqa: E402
from bench import WORKLOADS, make_graph, spmm_bytes  # noqa: E402
from laplace_gnn_recommendation_b200.dist import CudaOps  # noqa: E402


def locality_order(h, kind):
    """Row order = descending degree bucket (like lgb_degree_order), inside a bucket sorted by a column key of the row."""
    rp = h.rowptr.long()
    deg = rp[1:] - rp[:-1]
    n = deg.numel()
    lo, hi = int(rp[0]), int(rp[-1])
    cols = h.colidx[lo:hi].long()
    bucket = torch.where(deg <= 1, deg, 1 + torch.ceil(torch.log2(deg.clamp(min=2).double())).long())
    if kind == "firstcol":
        key = torch.where(deg > 0, h.colidx[(rp[:-1]).clamp(max=max(hi - 1, 0))].long(), torch.full_like(deg, h.n_cols))
    else:
        coldeg = torch.bincount(cols, minlength=h.n_cols)
        rows_of_e = torch.repeat_interleave(torch.arange(n, device=rp.device), deg)
        vals = coldeg[cols]
        m = torch.zeros(n, dtype=vals.dtype, device=rp.device).scatter_reduce_(0, rows_of_e, vals, "amax", include_self=True)
        idx = torch.nonzero(vals == m[rows_of_e]).view(-1)
        key = torch.full((n,), h.n_cols, dtype=torch.long, device=rp.device)
        key[rows_of_e[idx]] = cols[idx]
    o1 = torch.argsort(key, stable=True)
    o2 = torch.argsort(-bucket[o1], stable=True)
    return o1[o2].to(torch.int32).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--degree", default="powerlaw")
    ap.add_argument("--variants", default="1,3,6")
    ap.add_argument("--d", type=int, default=64)
    ap.add_argument("--workload", default="hm")
    ap.add_argument("--chunks", default="", help="slice sizes of the long-row plan to walk (default: the library default)")
    ap.add_argument("--rowsparse", type=int, default=0, help="B > 0: also time lgb_spmm_rowsparse on an operand that is zero outside the 3B rows of a BPR batch")
    ap.add_argument("--row-orders", default="", help="comma list of none|degree|firstcol|hotcol: processing order of the ordinary rows "
                    "(firstcol / hotcol: degree buckets, and inside a bucket rows sorted by their first / their most-referenced column)")
    ap.add_argument("--sweep", default="0", help="comma list of 0/1: long-row slices in plan order / column-sweep order")
    a = ap.parse_args()
    dev = _common.device()
    U, I, E = WORKLOADS[a.workload]
    if _common.DRYRUN:
        U, I, E = U // 400 + 8, I // 400 + 8, E // 400
    users, items = make_graph(U, I, E, a.degree, 1234, dev)
    row = torch.cat([users, items + U]); col = torch.cat([items + U, users])
    adj = lg.SparseTensor(row=row, col=col, sparse_sizes=(U + I, U + I))
    g = lg.gcn_norm(adj, add_self_loops=False).csr()
    ops = _common.make_ops(dev)
    gu, gi = ops.row_view(g, 0, U), ops.row_view(g, U, U + I)
    N, d = U + I, a.d
    X = torch.randn(N, d, device=dev); Y = torch.empty(N, d, device=dev); acc = torch.randn(N, d, device=dev)
    print(f"graph: nnz={g.nnz} n_long={g.n_long} n_tasks={g.n_tasks} | users view nnz={gu.nnz} long={gu.n_long} tasks={gu.n_tasks}"
          f" | items view nnz={gi.nnz} long={gi.n_long} tasks={gi.n_tasks}")
    Xs = bm = None
    if a.rowsparse > 0:          # a BPR batch: B edges drawn uniformly (user, positive item) + B uniform negatives
        gen = torch.Generator(device=dev).manual_seed(7)
        pick = torch.randint(0, users.numel(), (a.rowsparse,), generator=gen, device=dev)
        ub, pb = users[pick], items[pick]
        nb = torch.randint(0, I, (a.rowsparse,), generator=gen, device=dev)
        bm = lg.rows_bitmap(N, ((ub, 0), (pb, U), (nb, U)), dev)
        Xs = torch.zeros(N, d, device=dev)
        rows = torch.cat([ub, pb + U, nb + U])
        Xs[rows] = torch.randn(rows.numel(), d, device=dev)
        deg = (g.rowptr[1:] - g.rowptr[:-1]).long()
        print(f"rowsparse: {torch.unique(rows).numel()} operand rows flagged; entries that hit them: {int(deg[torch.unique(rows)].sum())} of {g.nnz}")
    chunks = [int(x) for x in a.chunks.split(",")] if a.chunks else [g.chunk]
    if a.row_orders:
        chunk, v = chunks[0], int(a.variants.split(",")[0])
        for h in (g, gu, gi):
            h._set_chunk(chunk)
            h.use_sweep_order(bool(int(a.sweep.split(",")[-1])))
        for kind in a.row_orders.split(","):
            for h in (g, gu, gi):
                if kind == "none":
                    h.use_degree_order(False)
                elif kind == "degree":
                    h.use_degree_order(False); h.use_degree_order(True)
                else:
                    h.row_order = locality_order(h, kind); h._struct = None
            ref = g.spmm(X, variant=0) if kind == a.row_orders.split(",")[0] else ref
            err = float((g.spmm(X, variant=v) - ref).abs().max())
            t_plain = timeit(lambda: g.spmm(X, Y=Y, variant=v))
            t_u = timeit(lambda: gu.spmm(X, Y=Y[:U], variant=v))
            t_i = timeit(lambda: gi.spmm(X, Y=Y[U:], variant=v))
            print(f"row order {kind:9s} chunk {chunk} variant {v}: plain {t_plain:.3f} ms | user rows {t_u:.3f} ms | item rows {t_i:.3f} ms | max diff vs first {err:.2e}", flush=True)
        return
    for chunk in chunks:
        for sweep in [bool(int(x)) for x in a.sweep.split(",")]:
            for h in (g, gu, gi):
                h._set_chunk(chunk)
                h.use_sweep_order(sweep)
            tag = f"chunk {chunk} sweep {int(sweep)} (item-view tasks {gi.n_tasks})"
            for v in [int(x) for x in a.variants.split(",")]:
                t_full = timeit(lambda: g.spmm(X, Y=Y, acc_in=acc, acc_out=acc, variant=v))
                t_plain = timeit(lambda: g.spmm(X, Y=Y, variant=v))
                t_u = timeit(lambda: gu.spmm(X, Y=Y[:U], variant=v))
                t_i = timeit(lambda: gi.spmm(X, Y=Y[U:], variant=v))
                gb = spmm_bytes(g.nnz, N, d) / 1e9
                print(f"{tag} variant {v}: full+acc {t_full:.3f} ms | plain {t_plain:.3f} ms ({gb / t_plain:.0f} GB/s alg) | "
                      f"user rows {t_u:.3f} ms ({spmm_bytes(gu.nnz, U, d) / 1e6 / t_u:.0f} GB/s) | "
                      f"item rows {t_i:.3f} ms ({spmm_bytes(gi.nnz, I, d) / 1e6 / t_i:.0f} GB/s)", flush=True)
            if bm is not None:
                t_d = timeit(lambda: g.spmm(Xs, Y=Y, resid=Xs))
                t_s = timeit(lambda: g.spmm(Xs, Y=Y, resid=Xs, x_rows=bm))
                t_su = timeit(lambda: gu.spmm(Xs, Y=Y[:U], x_rows=bm))
                t_si = timeit(lambda: gi.spmm(Xs, Y=Y[U:], x_rows=bm))
                Yd = g.spmm(Xs, resid=Xs); Ys = g.spmm(Xs, resid=Xs, x_rows=bm)
                err = float((Yd - Ys).abs().max()) / (float(Yd.abs().max()) + 1e-30)
                print(f"{tag} rowsparse B={a.rowsparse}: dense kernel on the sparse operand {t_d:.3f} ms | rowsparse {t_s:.3f} ms "
                      f"(user rows {t_su:.3f}, item rows {t_si:.3f}) | max rel diff {err:.2e}", flush=True)


if __name__ == "__main__":
    main()
