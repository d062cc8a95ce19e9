"""torchrun probe: per-rank component times of ShardedLightGCN (items SpMM, users SpMM, the item-block exchange -- own
kernel and NCCL --, forward and the full step under both schedules and as a CUDA graph).
torchrun --nproc-per-node N tools/dist_probe.py      PROBE_EXCHANGE=auto|symm|nccl  PROBE_GRAPH=1  PROBE_AUTOTUNE=1"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, make_graph  # noqa: E402
from laplace_gnn_recommendation_b200.dist import ShardedLightGCN  # noqa: E402


def timeit(fn, reps=20):
    for _ in range(5):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
    tmax, tmin = t.clone(), t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    return float(tmax), float(tmin)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    U, I, E = WORKLOADS["hm"]
    users, items = make_graph(U, I, E, "powerlaw", 1234, dev)
    eng = ShardedLightGCN(U, I, 64, 3, users, items, dev, exchange=os.environ.get("PROBE_EXCHANGE", "auto"))
    del users, items
    if os.environ.get("PROBE_AUTOTUNE", "1") == "1":
        eng.autotune()
    B = 128
    ub = torch.randint(0, U, (B,), device=dev); pb = torch.randint(0, I, (B,), device=dev); nb = torch.randint(0, I, (B,), device=dev)
    Ug = eng.Ug
    Xu, Xi, Yu, Yi = eng.table[:Ug], eng.table[Ug:], eng._yu[0], eng._yi[0]
    res = {}
    res["items_spmm"] = timeit(lambda: eng.ops.spmm(eng.g_items, Xu, Yi))
    res["users_spmm"] = timeit(lambda: eng.ops.spmm(eng.g_users, Xi, Yu, resid=eng._ru))
    res[f"exchange_items_block[{eng.ops.kind}{'/multicast' if getattr(eng.ops, 'multicast', False) else ''}]"] = \
        timeit(lambda: eng.ops.exchange_async(Yi, channel=0).wait())
    res["nccl_allreduce_items_block"] = timeit(lambda: dist.all_reduce(eng._ri))
    res["exchange_batch_rows"] = timeit(lambda: eng.ops.exchange_async(eng._stage_full.view(-1)[: 2 * B * 64], channel=2).wait())
    for sched in ("chains", "layer"):
        eng.schedule = sched
        res[f"forward[{sched}]"] = timeit(lambda: eng.forward())
        res[f"fused_step[{sched}]"] = timeit(lambda: eng.fused_step(ub, pb, nb, 1e-6))
    eng.schedule = "chains"
    if os.environ.get("PROBE_GRAPH", "1") == "1" and eng.ops.kind == "symm":
        gstep = eng.capture(B, 1e-6)
        res["fused_step[chains, CUDA graph]"] = timeit(lambda: gstep(ub, pb, nb))
    if dist.get_rank() == 0:
        print(f"world={dist.get_world_size()} Ug={eng.Ug} local_edges={eng.local_edges} "
              f"items view: nnz={eng.g_items.nnz} long={eng.g_items.n_long} tasks={eng.g_items.n_tasks}; "
              f"users view: nnz={eng.g_users.nnz} long={eng.g_users.n_long} tasks={eng.g_users.n_tasks}")
        for k, (a, b) in res.items():
            print(f"  {k:44s} max {a:.3f} ms   min {b:.3f} ms")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
