"""ONE-GPU probe of the per-rank SpMM launches of an N-way sharded run (no NCCL, no torchrun): builds rank r's local
block of the H&M-shaped graph exactly like dist.ShardedLightGCN does and times, per kernel variant and slice size,

    items SpMM (partial item rows <- owned users)   users SpMM (owned user rows <- items, fused epilogue)
    merged launch (lgb_spmm_split over all local rows)

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
    a = ap.parse_args()
    dev = _common.device()
    U, I, E = WORKLOADS["hm"]
    if _common.DRYRUN:
        U, I, E = U // 400, I // 400, E // 400
    users, items = make_graph(U, I, E, a.degree, 1234, dev)
    for rank in [int(x) for x in a.ranks.split(",")]:
        eng = ShardedLightGCN(U, I, a.d, 3, users, items, dev, rank=rank, world=a.world, ops=_common.make_ops(dev))   # no process group: local degrees
        g, Ug, n = eng.g_full, eng.Ug, eng.n
        X, Y, acc = eng.table, eng._ya, eng.E_f
        X.normal_(0, 0.1); acc.normal_(0, 0.1)
        print(f"rank {rank}/{a.world}: Ug={Ug} local_edges={eng.local_edges} max user deg={int((g.rowptr[1:Ug + 1] - g.rowptr[:Ug]).max())} "
              f"max item deg={int((g.rowptr[Ug + 1:] - g.rowptr[Ug:-1]).max())}")
        for chunk in [int(x) for x in a.chunks.split(",")]:
            view = lambda lo, hi: DeviceCSR(hi - lo, g.n_cols, g.rowptr[lo:hi + 1], g.colidx, g.val, chunk=chunk)   # noqa: E731
            gu, gi, ga = view(0, Ug), view(Ug, n), view(0, n)
            print(f"  chunk {chunk}: users view long={gu.n_long} tasks={gu.n_tasks} | items view long={gi.n_long} tasks={gi.n_tasks}")
            for v in [int(x) for x in a.variants.split(",")]:
                t_i = timeit(lambda: gi.spmm(X, Y=Y[Ug:], variant=v))
                t_u = timeit(lambda: gu.spmm(X, Y=Y[:Ug], acc_in=acc[:Ug], acc_out=acc[:Ug], variant=v))
                t_m = float("nan")
                if v not in (2, 3):
                    t_m = timeit(lambda: ga.spmm(X, Y=Y, acc_in=acc, acc_out=acc, variant=v, split_row=Ug, y_tail=Y[Ug:]))
                print(f"    variant {v:2d}: items {t_i * 1e3:7.1f} us | users {t_u * 1e3:7.1f} us | sum {(t_i + t_u) * 1e3:7.1f} us | "
                      f"merged launch {t_m * 1e3:7.1f} us   ({eng.local_edges / 1e6:.2f} M edges per direction)")
        del eng
        _common.empty_cache()


if __name__ == "__main__":
    main()
