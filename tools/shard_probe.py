"""ONE-GPU probe of the per-rank SpMM launches of an N-way sharded run (no NCCL, no torchrun): builds rank r's local
block of the H&M-shaped graph exactly like dist.ShardedLightGCN does and times, per kernel variant and slice size,

    items SpMM (partial item rows <- owned users)   users SpMM (owned user rows <- items, residual epilogue)
    both launches of a layer on two streams (the engine's "chains" schedule)

so the fixed per-launch cost seen in the scaling fit (profiles/README.md r1c: t = 0.09 ms + nnz / 34 G/s) can be attacked
at 1x GPU-minutes instead of 8x.      python tools/shard_probe.py [--world 8] [--ranks 0,3,7] [--variants 0,16,12]
                                                                  [--chunks 1024,256] [--degree powerlaw]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _common  # noqa: E402
from _common import timeit  # noqa: E402
from bench import WORKLOADS, make_graph  # noqa: E402
from laplace_gnn_recommendation_b200.csr import DeviceCSR  # noqa: E402
from laplace_gnn_recommendation_b200.dist import ShardedLightGCN  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--ranks", default="0,3,7")
    ap.add_argument("--variants", default="0,16,12")
    ap.add_argument("--chunks", default="1024,256")
    ap.add_argument("--degree", default="powerlaw")
    ap.add_argument("--d", type=int, default=64)
    ap.add_argument("--degree-orders", default="0", help="comma list of 0/1: ordinary rows in natural / degree-bucket order")
    a = ap.parse_args()
    dev = _common.device()
    U, I, E = WORKLOADS["hm"]
    if _common.DRYRUN:
        U, I, E = U // 400, I // 400, E // 400
    users, items = make_graph(U, I, E, a.degree, 1234, dev)
    for rank in [int(x) for x in a.ranks.split(",")]:
        eng = ShardedLightGCN(U, I, a.d, 3, users, items, dev, rank=rank, world=a.world, ops=_common.make_ops(dev))   # no process group: local degrees
        Ug, I = eng.Ug, eng.I
        gu0, gi0 = eng.g_users, eng.g_items
        Xu, Xi = eng.table[:Ug], eng.table[Ug:]
        Yu, Yi, R = torch.empty_like(Xu), torch.empty_like(Xi), torch.randn_like(Xu)
        eng.table.normal_(0, 0.1)
        print(f"rank {rank}/{a.world}: Ug={Ug} local_edges={eng.local_edges} max user deg={int((gu0.rowptr[1:] - gu0.rowptr[:-1]).max())} "
              f"max item deg={int((gi0.rowptr[1:] - gi0.rowptr[:-1]).max())}")
        s0, s1 = (None, None) if _common.DRYRUN else (torch.cuda.Stream(), torch.cuda.Stream())

        def both(gu, gi, v):                       # the two SpMM launches of one layer on two streams (the engine's "chains" form)
            cur = torch.cuda.current_stream()
            s0.wait_stream(cur); s1.wait_stream(cur)
            with torch.cuda.stream(s0):
                gi.spmm(Xu, Y=Yi, variant=v)
            with torch.cuda.stream(s1):
                gu.spmm(Xi, Y=Yu, resid=R, variant=v)
            cur.wait_stream(s0); cur.wait_stream(s1)
        for chunk in [int(x) for x in a.chunks.split(",")]:
            def view(g0):
                v = DeviceCSR(g0.n_rows, g0.n_cols, g0.rowptr, g0.colidx, g0.val, chunk=chunk)
                v.nnz = g0.nnz
                return v
            gu, gi = view(gu0), view(gi0)
            print(f"  chunk {chunk}: users view long={gu.n_long} tasks={gu.n_tasks} | items view long={gi.n_long} tasks={gi.n_tasks}")
            for v, order in [(int(x), bool(int(o))) for o in a.degree_orders.split(",") for x in a.variants.split(",")]:
                gu.use_degree_order(order); gi.use_degree_order(order)
                if len(a.degree_orders.split(",")) > 1:
                    print(f"    degree order {int(order)}:", end="")
                t_i = timeit(lambda: gi.spmm(Xu, Y=Yi, variant=v))
                t_u = timeit(lambda: gu.spmm(Xi, Y=Yu, resid=R, variant=v))
                t_b = timeit(lambda: both(gu, gi, v)) if not _common.DRYRUN else float("nan")
                print(f"    variant {v:2d}: items {t_i * 1e3:7.1f} us | users {t_u * 1e3:7.1f} us | sum {(t_i + t_u) * 1e3:7.1f} us | "
                      f"both on two streams {t_b * 1e3:7.1f} us   ({eng.local_edges / 1e6:.2f} M edges per direction)")
        del eng
        _common.empty_cache()


if __name__ == "__main__":
    main()
