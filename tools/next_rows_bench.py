"""Measurement of the SURVEY 8f "next" rows at the H&M shape, one JSON line each (CUDA events, roofline per row):

  f3  fused Adam step         7*N*d*4 bytes per step (p, g, m, v read; p, m, v written)      vs the HBM copy peak
  f2  score + top-k           users/s; L2 bytes = users * I*d*4 (the item table is re-read per user)
  f1  negative-sample rejection (device part of structured_negative_sampling): sorted-key build + binary-search mask
  L9  bpr_loss over every edge of a split (evaluation, B = E_split, forward only)              B*6*d*4 bytes

    python tools/next_rows_bench.py [--rows adam,topk,sampler,bpr]        LGB_TOOLS_DRYRUN=1 for a logic check without a GPU
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _common  # noqa: E402
from _common import timeit  # noqa: E402
import laplace_gnn_recommendation_b200 as lg  # noqa: E402
from bench import WORKLOADS, make_graph  # noqa: E402


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(_common.REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", default="adam,topk,sampler,bpr")
    ap.add_argument("--d", type=int, default=64)
    a = ap.parse_args()
    dev = _common.device()
    U, I, E = WORKLOADS["hm"]
    topk_users = 4096
    if _common.DRYRUN:
        U, I, E, topk_users = 400, 150, 6000, 32
    N, d, peak = U + I, a.d, peak_gbs()
    rows = a.rows.split(",")

    if "adam" in rows:
        p = [torch.randn(U, d, device=dev).requires_grad_(True), torch.randn(I, d, device=dev).requires_grad_(True)]
        for t in p:
            t.grad = torch.randn_like(t)
        fused, stock = lg.FusedAdam(p, lr=1e-3), torch.optim.Adam([t.detach().clone().requires_grad_(True) for t in p], lr=1e-3)
        for t, g in zip(stock.param_groups[0]["params"], p):
            t.grad = g.grad.clone()
        ms_f, ms_t = timeit(fused.step), timeit(stock.step)
        nbytes = 7 * N * d * 4
        print(json.dumps({"row": "f3 fused Adam", "ms": ms_f, "GBps": nbytes / 1e6 / ms_f, "frac_of_hbm_peak": nbytes / 1e6 / ms_f / peak,
                          "algorithmic_bytes": nbytes, "torch_optim_adam_ms": ms_t, "speedup_vs_torch": ms_t / ms_f}), flush=True)

    if "topk" in rows:
        Wu, Wi = torch.randn(U, d, device=dev) * 0.1, torch.randn(I, d, device=dev) * 0.1
        users, items = make_graph(U, I, E, "powerlaw", 1234, dev)
        seen = lg.SeenItems(torch.stack([users, items]), U, I)
        who = torch.randint(0, U, (topk_users,), device=dev)
        for k, mode in ((12, "cta"), (12, "tiled"), (256, "cta"), (256, "tiled")):
            kk = min(k, I)
            ms = timeit(lambda: lg.recommend_topk(Wu, Wi, who, kk, seen, mode=mode), reps=5, warmup=2)
            print(json.dumps({"row": "f2 score + top-k with exclusion", "scoring": mode, "k": kk, "users": topk_users, "ms": ms,
                              "fp32_TFLOPs": 2.0 * topk_users * I * d / 1e9 / ms,
                              "users_per_s": topk_users / (ms * 1e-3), "item_table_GBps": topk_users * I * d * 4 / 1e6 / ms,
                              "all_users_estimate_s": U / (topk_users / (ms * 1e-3))}), flush=True)

    if "sampler" in rows or "bpr" in rows:
        users, items = make_graph(U, I, E, "powerlaw", 1234, dev)
        ei = torch.stack([users, items])

    if "sampler" in rows:
        num_nodes = int(ei[1].max())
        torch.manual_seed(0)
        t0 = time.perf_counter()
        lg.structured_negative_sampling(ei, num_nodes=num_nodes)        # first call: builds + caches the sorted key array
        _common.sync(); first = time.perf_counter() - t0
        t0 = time.perf_counter()
        reps = 1 if _common.DRYRUN else 3
        for _ in range(reps):
            lg.structured_negative_sampling(ei, num_nodes=num_nodes)
        _common.sync(); steady = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        torch.randint(num_nodes, (E,), dtype=torch.long)
        cpu_rng = time.perf_counter() - t0
        print(json.dumps({"row": "f1 structured_negative_sampling (bit-exact CPU draws + device rejection)", "E": E, "first_call_s": first,
                          "steady_s": steady, "of_which_cpu_randint_s": cpu_rng, "edges_per_s": E / steady}), flush=True)

    if "bpr" in rows:
        Ef, E0 = torch.randn(N, d, device=dev) * 0.1, torch.randn(N, d, device=dev) * 0.1
        neg = torch.randint(0, I, (E,), device=dev)
        loss = torch.empty((), device=dev)
        ms = timeit(lambda: lg.bpr_indexed(Ef, E0, U, users, items, neg, 1e-6, loss=loss), reps=10, warmup=3)
        nbytes = E * (6 * d * 4 + 24)
        print(json.dumps({"row": "L9 bpr_loss over every edge (evaluation)", "B": E, "ms": ms, "GBps": nbytes / 1e6 / ms,
                          "frac_of_hbm_peak": nbytes / 1e6 / ms / peak, "algorithmic_bytes": nbytes}), flush=True)


if __name__ == "__main__":
    main()
