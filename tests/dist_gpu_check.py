"""torchrun helper (not collected by pytest): N ranks over NCCL, each checks its shard of ShardedLightGCN
against the single-process CPU oracle.  Prints DIST_OK per rank."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laplace_gnn_recommendation_b200.dist import ShardedLightGCN  # noqa: E402
from tests.test_dist_gloo import make_problem, single_process_reference  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    # DIST_CHECK_MODE = "schedule,exchange[,graph]", e.g. "chains,symm,graph"; default = what bench.py ships
    mode = os.environ.get("DIST_CHECK_MODE", "chains,auto").split(",")
    if mode[0] == "rows":                      # DIST_CHECK_MODE = "rows,R" / "rows,S": the generic row-sharded engine
        return check_rows(dev, mode[1] if len(mode) > 1 else "R")
    for K in (3, 2):
        pb = make_problem(seed=K, U=2000, I=300, E=40000, d=64, K=K, B=512)
        eng = ShardedLightGCN(pb["U"], pb["I"], pb["d"], K, pb["users"], pb["items"], dev,
                              init_tables=(pb["Wu"].to(dev), pb["Wi"].to(dev)),
                              schedule=mode[0], exchange=mode[1], max_batch=512)
        if os.environ.get("DIST_CHECK_ROWSPARSE") == "1":      # first backward layer as lgb_spmm_rowsparse whatever the sizes
            eng.rowsparse_ratio = 0
        step = eng.capture(512, pb["lam"]) if "graph" in mode[2:] else (lambda u, p, n: eng.fused_step(u, p, n, pb["lam"]))
        for _ in range(3):                    # repeated: every buffer, signal slot and stream dependency is re-used
            loss = step(pb["u"].to(dev), pb["p"].to(dev), pb["n"].to(dev))
        torch.cuda.synchronize()
        o_loss, o_gu, o_gi, o_uf, o_if = single_process_reference(pb)
        tol = dict(rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(loss.cpu(), o_loss, **tol)
        torch.testing.assert_close(eng.E_f[: eng.Ug].cpu(), o_uf[eng.lo:eng.hi], **tol)
        torch.testing.assert_close(eng.E_f[eng.Ug:].cpu(), o_if, **tol)
        torch.testing.assert_close(eng.grad[: eng.Ug].cpu(), o_gu[eng.lo:eng.hi], rtol=1e-5, atol=1e-9)
        torch.testing.assert_close(eng.grad[eng.Ug:].cpu(), o_gi, rtol=1e-5, atol=1e-9)
    print(f"DIST_OK rank={dist.get_rank()} mode={','.join(mode)} rowsparse={os.environ.get('DIST_CHECK_ROWSPARSE', '0')} exchange={eng.ops.kind} multicast={getattr(eng.ops, 'multicast', False)} "
          f"users=[{eng.lo},{eng.hi}) edges={eng.local_edges}", flush=True)
    dist.destroy_process_group()


def check_rows(dev, wiring):
    from laplace_gnn_recommendation_b200.dist_rows import RowShardedLightGCN
    from oracle import lightgcn_oracle as lo
    for K in (3, 1):
        pb = make_problem(seed=K, U=2000, I=300, E=40000, d=64, K=K, B=512)
        row, col, n = (lo.wiring_reference if wiring == "R" else lo.wiring_symmetric)(pb["users"], pb["items"], pb["U"], pb["I"])
        eng = RowShardedLightGCN(pb["U"], pb["I"], pb["d"], K, row, col, dev, init_tables=(pb["Wu"], pb["Wi"]))
        loss = eng.fused_step(pb["u"], pb["p"], pb["n"], pb["lam"])
        torch.cuda.synchronize()
        rowptr, c, _ = lo.csr_from_coo(row, col, n, n)
        o_loss, o_gu, o_gi, o_uf, o_if = lo.train_iteration(pb["Wu"], pb["Wi"], rowptr, c, K, pb["u"], pb["p"], pb["n"], pb["lam"])
        o_ef, o_g = torch.cat([o_uf, o_if]), torch.cat([o_gu, o_gi])
        torch.testing.assert_close(loss.cpu(), o_loss, rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(eng.E_f.cpu(), o_ef[eng.lo:eng.hi], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(eng.grad.cpu(), o_g[eng.lo:eng.hi], rtol=1e-5, atol=1e-9)
    print(f"DIST_OK rank={dist.get_rank()} rows engine wiring={wiring} nodes=[{eng.lo},{eng.hi}) nnz={eng.local_nnz}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
