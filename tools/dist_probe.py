"""torchrun probe: per-rank component times of one ShardedLightGCN layer (items SpMM, users SpMM, all-reduce of
the item block, full layer, full fused_step).  torchrun --nproc-per-node N tools/dist_probe.py"""
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
    eng = ShardedLightGCN(U, I, 64, 3, users, items, dev)
    del users, items
    B = 128
    ub = torch.randint(0, U, (B,), device=dev); pb = torch.randint(0, I, (B,), device=dev); nb = torch.randint(0, I, (B,), device=dev)
    X, Y, Ug = eng.table, eng._ya, eng.Ug
    res = {}
    res["items_spmm"] = timeit(lambda: eng.ops.spmm(eng.g_items, X, Y=Y[Ug:]))
    res["users_spmm"] = timeit(lambda: eng.ops.spmm(eng.g_users, X, Y=Y[:Ug], acc_in=eng.E_f[:Ug], acc_out=eng.E_f[:Ug]))
    res["allreduce_items_block"] = timeit(lambda: dist.all_reduce(Y[Ug:]))
    res["accumulate_items"] = timeit(lambda: eng.ops.accumulate(Y[Ug:], eng.E_f[Ug:], None, 1.0, eng.E_f[Ug:]))
    res["forward"] = timeit(lambda: eng.forward())
    res["fused_step"] = timeit(lambda: eng.fused_step(ub, pb, nb, 1e-6))
    if os.environ.get("PROBE_PIPELINED"):
        eng.schedule = "pipelined"
        res["fused_step_pipelined"] = timeit(lambda: eng.fused_step(ub, pb, nb, 1e-6))
        eng.static_batch = True
        res["fused_step_pipelined_static"] = timeit(lambda: eng.fused_step(ub, pb, nb, 1e-6))
    if os.environ.get("PROBE_GRAPH"):
        gstep = eng.capture(B, 1e-6)
        res["fused_step_graph"] = timeit(lambda: gstep(ub, pb, nb))
    if dist.get_rank() == 0:
        print(f"world={dist.get_world_size()} Ug={eng.Ug} local_edges={eng.local_edges} "
              f"items view: nnz={eng.g_items.nnz} long={eng.g_items.n_long} tasks={eng.g_items.n_tasks}; "
              f"users view: nnz={eng.g_users.nnz} long={eng.g_users.n_long} tasks={eng.g_users.n_tasks}")
        for k, (a, b) in res.items():
            print(f"  {k:24s} max {a:.3f} ms   min {b:.3f} ms")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
