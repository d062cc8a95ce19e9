"""The reference's LightGCN driver loop (run_pipeline_lightgcn.py:76-222 `train`) on this package -- the drop-in shown end to
end on a synthetic graph: create_dataloaders_lightgcn-style splits, LightGCN, Adam + ExponentialLR, per iteration
sample_mini_batch -> forward -> six gathers -> bpr_loss -> backward -> step, evaluation() every `--eval-every`, top-k dump.

    python tools/train_lightgcn.py [--workload ml1m] [--iters 200] [--style reference|fused] [--batch 128]

`--style reference` keeps the reference's own call sequence line for line (only the imports differ); `--style fused` swaps
the iteration body for `model.fused_step` + `FusedAdam`.  Prints one JSON line with iterations/s split into sampler / step /
evaluation time, so the CPU-side sampler cost the reference pays per iteration is visible next to the GPU part.
LGB_TOOLS_DRYRUN=1 runs a tiny graph on the CPU emulation of the kernels (logic check, no GPU).
"""
import argparse
import json
import os
import random
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _common  # noqa: E402
import laplace_gnn_recommendation_b200 as lg  # noqa: E402
from bench import WORKLOADS, make_graph  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ml1m", choices=sorted(WORKLOADS))
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--style", default="reference", choices=["reference", "fused", "both"])
    ap.add_argument("--batch", type=int, default=128)          # config.py:144
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--lam", type=float, default=1e-6)
    ap.add_argument("--k", type=int, default=12)
    ap.add_argument("--eval-every", type=int, default=100)
    ap.add_argument("--lr-decay-every", type=int, default=50)
    ap.add_argument("--sampler", default="reference", choices=["reference", "device"],
                    help="reference: bit-exact sample_mini_batch (O(E) CPU draws per iteration); device: DeviceSampler (O(B), same distribution)")
    ap.add_argument("--autotune", action="store_true", help="plan-time kernel / plan selection on the training graph (LightGCN.autotune)")
    a = ap.parse_args()
    for style in (["reference", "fused"] if a.style == "both" else [a.style]):
        a.style = style
        run(a)


def run(a):
    dev = _common.device()
    U, I, E = WORKLOADS[a.workload]
    if _common.DRYRUN:
        U, I, E, a.iters, a.eval_every, a.lr_decay_every = 60, 40, 900, 6, 3, 2
    torch.manual_seed(0); random.seed(0); np.random.seed(0)

    # --- create_dataloaders_lightgcn (data/lightgcn_loader.py:54-91) on an in-memory homogeneous edge list
    users, items = make_graph(U, I, E, "powerlaw", 1234, torch.device("cpu"))
    keys = torch.unique(users * I + items)                    # distinct interactions, like the transactions table
    homo = torch.stack([keys // I, keys % I + U])             # to_homogeneous(): item ids follow the user ids
    (train_sp, val_sp, test_sp, train_ei, val_ei, test_ei, edge_index, num_users, num_items) = lg.make_lightgcn_splits(homo, U, I)

    model = lg.LightGCN(num_users, num_items, embedding_dim=a.dim, num_iterations=a.layers).to(dev)
    model.train()
    opt = (lg.FusedAdam if a.style == "fused" else torch.optim.Adam)(model.parameters(), lr=a.lr)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.95)
    train_ei, val_ei, test_ei = train_ei.to(dev), val_ei.to(dev), test_ei.to(dev)
    train_sp, val_sp, test_sp = train_sp.to(dev), val_sp.to(dev), test_sp.to(dev)

    t_sample = t_step = t_eval = 0.0
    log = []
    t_tune, plan = 0.0, None
    if getattr(a, "autotune", False):
        _common.sync(); t0 = time.perf_counter()
        model.autotune(train_sp)
        _common.sync(); t_tune = time.perf_counter() - t0
        g = lg.gcn_norm(train_sp, add_self_loops=False).csr()
        plan = {"forward": g.autotune_report.get("chosen"), "backward": g.transpose().autotune_report.get("chosen")}
    dsampler = lg.DeviceSampler(train_ei) if a.sampler == "device" else None
    for it in range(a.iters):
        t0 = time.perf_counter()
        if dsampler is not None:
            u_idx, p_idx, n_idx = dsampler.sample(a.batch)
        else:
            u_idx, p_idx, n_idx = lg.sample_mini_batch(a.batch, train_ei)
            u_idx, p_idx, n_idx = u_idx.to(dev), p_idx.to(dev), n_idx.to(dev)
        _common.sync(); t1 = time.perf_counter()
        if a.style == "reference":          # run_pipeline_lightgcn.py:120-158, verbatim call sequence
            users_emb_final, users_emb_0, items_emb_final, items_emb_0 = model.forward(train_sp)
            train_loss = lg.bpr_loss(users_emb_final[u_idx], users_emb_0[u_idx], items_emb_final[p_idx], items_emb_0[p_idx],
                                     items_emb_final[n_idx], items_emb_0[n_idx], a.lam)
            opt.zero_grad()
            train_loss.backward()
            opt.step()
        else:
            train_loss = model.fused_step(train_sp, u_idx, p_idx, n_idx, a.lam)
            opt.step()
        _common.sync(); t2 = time.perf_counter()
        t_sample += t1 - t0; t_step += t2 - t1
        if it % a.eval_every == 0:
            model.eval()
            val_loss, recall, precision, ndcg = lg.evaluation(model, val_ei, val_sp, [train_ei], a.k, a.lam)
            model.train()
            log.append(dict(iter=it, train_loss=round(float(train_loss), 6), val_loss=round(val_loss, 6), recall=round(recall, 6),
                            precision=round(precision, 6), ndcg=round(ndcg, 6)))
            _common.sync(); t_eval += time.perf_counter() - t2
        if it % a.lr_decay_every == 0 and it != 0:
            sched.step()
    model.eval()
    test = lg.evaluation(model, test_ei, test_sp, [train_ei, val_ei], a.k, a.lam)
    # final candidate dump (run_pipeline_lightgcn.py:211-222): top-k per user from the layer-0 tables, seen items excluded
    seen = lg.SeenItems(edge_index.to(dev), num_users, num_items)
    cands = lg.recommend_topk(model.users_emb.weight.detach(), model.items_emb.weight.detach(),
                              torch.arange(min(num_users, 1024), device=dev), min(256, num_items), seen)
    print(json.dumps({"workload": a.workload if not _common.DRYRUN else "dryrun", "style": a.style, "sampler": a.sampler, "iters": a.iters,
                      "iters_per_s": a.iters / max(t_sample + t_step, 1e-9), "sampler_s": t_sample, "step_s": t_step, "eval_s": t_eval,
                      "autotune_s": t_tune, "plan": plan,
                      "test": dict(zip(("loss", "recall", "precision", "ndcg"), [round(float(x), 6) for x in test])),
                      "candidates_shape": list(cands.shape), "log": log}), flush=True)


if __name__ == "__main__":
    main()
