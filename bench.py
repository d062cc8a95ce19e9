#!/usr/bin/env python
"""bench.py -- LightGCN fwd+bwd+BPR epoch throughput (edge-traversals/s) on synthetic H&M-shaped graphs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload hm|ml1m]

A "step" = one epoch in the sense of SURVEY.md 8d = one pass of run_pipeline_lightgcn.py:118-158: full-graph
forward (K SpMM + layer mean), BPR over B sampled triples, full backward (K SpMM^T) -- without the CPU
sampler and the Adam step.  edge-traversals/s = 2*K*nnz(A_hat)/t_epoch.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import torch  # noqa: E402

WORKLOADS = {
    # name: (U, I, E)  -- BASELINE.json configs[2] (H&M-shaped) and configs[1] (MovieLens-1M-shaped)
    "hm": (1_371_980, 105_542, 31_788_324),
    "ml1m": (6_040, 3_706, 1_000_209),
}
CPU_SAMPLE_SCALE = 8  # the CPU arms run the same generator at 1/8 of every dimension (bounded sample)


def zipf_ids(n_ids: int, n_samples: int, alpha: float, gen: torch.Generator, device) -> torch.Tensor:
    """Ids drawn from Zipf(alpha) over a random permutation of [0, n_ids) (SURVEY 8d degree model)."""
    ranks = torch.arange(1, n_ids + 1, dtype=torch.float64, device=device)
    cdf = torch.cumsum(ranks.pow(-alpha), 0)
    cdf = cdf / cdf[-1]
    u = torch.rand(n_samples, generator=gen, device=device, dtype=torch.float64)
    r = torch.searchsorted(cdf, u).clamp_(max=n_ids - 1)
    perm = torch.randperm(n_ids, generator=gen, device=device)
    return perm[r]


def make_graph(U: int, I: int, E: int, degree: str, seed: int, device):
    gen = torch.Generator(device=device).manual_seed(seed)
    if degree == "powerlaw":
        users = zipf_ids(U, E, 0.8, gen, device)
        items = zipf_ids(I, E, 1.1, gen, device)
    else:
        users = torch.randint(0, U, (E,), generator=gen, device=device)
        items = torch.randint(0, I, (E,), generator=gen, device=device)
    return users, items


def spmm_bytes(nnz: int, rows: int, d: int) -> float:
    """Algorithmic bytes of one SpMM launch, cache-oblivious row-gather model."""
    return nnz * 8 + (rows + 1) * 4 + nnz * d * 4 + rows * d * 4


def epoch_bytes(nnz: int, N: int, d: int, K: int, B: int) -> float:
    bpr = B * 12 * d * 4 + 3 * B * 8 + N * d * 4
    return 2 * K * spmm_bytes(nnz, N, d) + (3 * K + 2) * N * d * 4 + bpr


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_epoch_runner(U, I, E, d, K, B, degree, seed):
    """The CPU arm: the oracle's training iteration (forward, gathers, bpr_loss, autograd backward) with the
    SpMM served by torch's multi-threaded CSR matmul -- the stand-in for torch_sparse's CPU spmm."""
    from oracle import lightgcn_oracle as lo
    users, items = make_graph(U, I, E, degree, seed, torch.device("cpu"))
    row, col, n = lo.wiring_symmetric(users, items, U, I)
    rowptr, c, _ = lo.csr_from_coo(row, col, n, n)
    _, val = lo.gcn_norm_values(rowptr, c)
    mm = lo.make_cpu_matmul(rowptr, c, val, n)
    gen = torch.Generator().manual_seed(0)
    Wu, Wi = torch.randn(U, d, generator=gen) * 0.1, torch.randn(I, d, generator=gen) * 0.1
    pick = torch.randint(0, E, (B,), generator=gen)
    ub, pb, nb = users[pick], items[pick], torch.randint(0, I, (B,), generator=gen)

    def step():
        return lo.train_iteration(Wu, Wi, rowptr, c, K, ub, pb, nb, 1e-6, val=val, matmul=mm)[0]
    return step, 2 * E


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the reference itself cannot be imported:
    torch_sparse / torch_geometric are not installable here) on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    U, I, E = (x // CPU_SAMPLE_SCALE for x in WORKLOADS[args.workload])
    step, nnz = cpu_epoch_runner(U, I, E, args.dim, args.layers, args.batch, args.degree, 1234)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = 2 * args.layers * nnz / dt
    sample = (f"{args.workload}-shaped graph at 1/{CPU_SAMPLE_SCALE} scale (U={U}, I={I}, E={E}, nnz={nnz}, {args.degree}), "
              f"full epoch per step, torch CSR (MKL) SpMM + autograd")
    line = base_line(args, value, dt * 1e3, nnz_full=None)
    line.update({"impl": "reference", "n_gpus": args.gpus, "dtype": "f32",
                 "cpu_baseline": {"value": value, "unit": "edge-traversals/s", "cores": torch.get_num_threads(),
                                  "kind": "port", "sample": sample},
                 "e2e": {"value": value, "unit": "edge-traversals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)


def base_line(args, value, ms, nnz_full):
    U, I, E = WORKLOADS[args.workload]
    return {
        "metric": "lightgcn_fwd_bwd_bpr_edge_traversals_per_s", "value": value, "unit": "edge-traversals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}_{args.degree}_symmetric", "U": U, "I": I, "E": E, "nnz": 2 * E,
                   "d": args.dim, "K": args.layers, "B": args.batch,
                   "epoch": "K SpMM fwd + layer mean + BPR(B) + K SpMM^T bwd; sampler and Adam excluded",
                   "l2": "working set (tables 378 MB + CSR 1 GB per direction at hm) exceeds the 126 MB L2; no flush",
                   "parallelism": "single GPU" if args.gpus == 1 else f"users range-sharded x{args.gpus}, items replicated, NCCL all-reduce"},
    }


def run_ours(args):
    import torch.distributed as dist

    import laplace_gnn_recommendation_b200 as lg
    from laplace_gnn_recommendation_b200 import _lib
    from laplace_gnn_recommendation_b200.csr import DeviceCSR

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")

    U, I, E = WORKLOADS[args.workload]
    d, K, B, lam = args.dim, args.layers, args.batch, 1e-6
    tuned = None
    users, items = make_graph(U, I, E, args.degree, 1234, dev)
    gen = torch.Generator(device=dev).manual_seed(42)
    pick = torch.randint(0, E, (B,), generator=gen, device=dev)
    ub, pb = users[pick].contiguous(), items[pick].contiguous()
    nb = torch.randint(0, I, (B,), generator=gen, device=dev)

    if world == 1:
        row = torch.cat([users, items + U]); col = torch.cat([items + U, users])
        adj = lg.SparseTensor(row=row, col=col, sparse_sizes=(U + I, U + I))
        del row, col
        torch.manual_seed(0)
        model = lg.LightGCN(U, I, d, K).to(dev)
        g = lg.gcn_norm(adj, add_self_loops=False).csr()
        g.transpose()
        if args.degree_order:
            g.use_degree_order(); g.transpose().use_degree_order()
        if not args.no_autotune:     # plan-time choice of the SpMM kernel variant for this graph (result-checked, see csr.py)
            gt = g.transpose()
            try:
                g.autotune(d, chunks=(1024, 512), degree_orders=(False, True))
                gt.autotune(d, chunks=(1024, 512), degree_orders=(False, True))
                tuned = {"forward": g.autotune_report.get("chosen"), "backward": gt.autotune_report.get("chosen"),
                         "forward_ms": g.autotune_report["ms"], "backward_ms": gt.autotune_report["ms"],
                         "rejected": {**g.autotune_report["rejected"], **gt.autotune_report["rejected"]}}
            except Exception as exc:      # tuning is optional: fall back to the default plan and say so
                g.variant = gt.variant = None
                tuned = {"error": repr(exc)[:300]}
        nnz = g.nnz
        step = lambda: model.fused_step(adj, ub, pb, nb, lam)   # noqa: E731
    else:
        from laplace_gnn_recommendation_b200.dist import ShardedLightGCN
        torch.manual_seed(0)
        eng = ShardedLightGCN(U, I, d, K, users, items, dev, schedule=args.schedule, exchange=args.exchange)
        if not args.no_autotune:
            try:
                tuned = eng.autotune()                        # per rank; rank 0's choice is reported
            except Exception as exc:
                for gv in eng.graphs():
                    gv.variant = None
                tuned = {"error": repr(exc)[:300]}
        nnz = 2 * E
        if not args.graph:
            step = lambda: eng.fused_step(ub, pb, nb, lam)      # noqa: E731
        else:                                                   # whole step (kernels + NCCL) replayed from one CUDA graph
            gstep = eng.capture(B, lam)
            step = lambda: gstep(ub, pb, nb)                    # noqa: E731
    del users, items
    torch.cuda.empty_cache()

    # per-launch CUDA-event timing of the dominant kernel (lgb_spmm) inside the timed region
    spmm_events = []
    orig_spmm = DeviceCSR.spmm

    def timed_spmm(self, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_spmm(self, *a, **k)
        e1.record()
        spmm_events.append((e0, e1, self.nnz, self.n_rows))
        return out

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    DeviceCSR.spmm = timed_spmm
    launches0 = _lib.LAUNCHES
    sync()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        loss = step()
    t1.record()
    sync()
    launches = _lib.LAUNCHES - launches0
    clk = clocks.stop() if rank == 0 else None
    graphed = world > 1 and args.graph
    if graphed:
        # the timed region replayed a CUDA graph (no per-launch host hooks); per-launch SpMM durations for the
        # roofline and the launch count come from an instrumented kernel-by-kernel pass of the SAME step
        launches0 = _lib.LAUNCHES
        for _ in range(3):
            eng.fused_step(ub, pb, nb, lam)
        launches = (_lib.LAUNCHES - launches0) // 3 * args.steps
        sync()
    DeviceCSR.spmm = orig_spmm
    ms = torch.tensor([t0.elapsed_time(t1) / max(args.steps, 1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    value = 2 * K * nnz / (ms * 1e-3)

    spmm_ms = [a.elapsed_time(b) for a, b, _, _ in spmm_events]
    spmm_alg = [spmm_bytes(z, r, d) for _, _, z, r in spmm_events]
    mp = {}
    try:
        mp = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(mp.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in mp else "fallback 6650 GB/s (B200_PROFILING.md)"
    achieved = (sum(spmm_alg) / 1e9) / (sum(spmm_ms) * 1e-3) if spmm_ms else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(REPO, "profiles", "roofline_traffic.json"))).get(args.workload)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "lgb_spmm (lgb::spmm_subwarp_kernel / spmm_rows_kernel, variant per config.spmm_variant)",
                "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": spmm_alg[0] if spmm_alg else None,
                "avg_launch_ms": statistics.mean(spmm_ms) if spmm_ms else None,
                "launches_timed": len(spmm_ms),
                "timed_in": "separate kernel-by-kernel pass (timed region replays a CUDA graph)" if graphed else "timed region",
                "spmm_share_of_step": (sum(spmm_ms) / ((3 if graphed else args.steps) * ms)) if spmm_ms else None,
                "epoch_algorithmic_gb": epoch_bytes(nnz, U + I, d, K, B) / 1e9,
                "epoch_frac_of_peak": epoch_bytes(nnz, U + I, d, K, B) / 1e9 / (ms * 1e-3) / peak}

    # ---- e2e: the reference-style call sequence through the public API, host buffers in the timed region
    e2e = None
    if world == 1:
        hu, hp, hn = (t.cpu().pin_memory() for t in (ub, pb, nb))

        def api_step():
            u_ = hu.to(dev, non_blocking=True); p_ = hp.to(dev, non_blocking=True); n_ = hn.to(dev, non_blocking=True)
            u_f, u_0, i_f, i_0 = model(adj)
            l = lg.bpr_loss(u_f[u_], u_0[u_], i_f[p_], i_0[p_], i_f[n_], i_0[n_], lam)
            model.zero_grad(set_to_none=True)
            l.backward()
            return l.item()                       # device -> host read of the step's result
        for _ in range(max(3, args.warmup // 2)):
            api_step()
        sync()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            api_step()
        torch.cuda.synchronize()
        e_ms = (time.perf_counter() - w0) * 1e3 / max(args.steps, 1)
        e2e = {"value": 2 * K * nnz / (e_ms * 1e-3), "unit": "edge-traversals/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": 3 * B * 8, "d2h_bytes_per_step": 4,
               "path": "LightGCN.forward(SparseTensor) -> 6 gathers -> bpr_loss -> loss.backward() -> loss.item()"}
    else:
        hu, hp, hn = (t.cpu().pin_memory() for t in (ub, pb, nb))

        def api_step():
            if not args.graph:
                u_ = hu.to(dev, non_blocking=True); p_ = hp.to(dev, non_blocking=True); n_ = hn.to(dev, non_blocking=True)
            return (eng.fused_step(u_, p_, n_, lam) if not args.graph else gstep(hu, hp, hn)).item()
        for _ in range(3):
            api_step()
        sync()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            api_step()
        sync()
        e_ms = torch.tensor([(time.perf_counter() - w0) * 1e3 / max(args.steps, 1)], device=dev, dtype=torch.float64)
        dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
        e2e = {"value": 2 * K * nnz / (float(e_ms) * 1e-3), "unit": "edge-traversals/s", "ms_per_step": float(e_ms),
               "h2d_bytes_per_step": 3 * B * 8, "d2h_bytes_per_step": 4,
               "path": "ShardedLightGCN.fused_step(host-pinned batch) -> loss.item()"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        Us, Is, Es = (x // CPU_SAMPLE_SCALE for x in WORKLOADS[args.workload])
        cstep, cnnz = cpu_epoch_runner(Us, Is, Es, d, K, B, args.degree, 1234)
        cstep()
        c0 = time.perf_counter()
        reps = 0
        while reps < 3 or (time.perf_counter() - c0 < 10 and reps < 20):
            cstep(); reps += 1
        cdt = (time.perf_counter() - c0) / reps
        cpu_baseline = {"value": 2 * K * cnnz / cdt, "unit": "edge-traversals/s", "cores": torch.get_num_threads(),
                        "kind": "port", "ms_per_step": cdt * 1e3,
                        "sample": f"{args.workload}-shaped graph at 1/{CPU_SAMPLE_SCALE} scale (U={Us}, I={Is}, E={Es}, nnz={cnnz}, "
                                  f"{args.degree}), {reps} full epochs after 1 warm-up; oracle port, torch CSR (MKL) SpMM + autograd"}

    if rank == 0:
        line = base_line(args, value, ms, nnz)
        line["config"]["spmm_variant"] = tuned if tuned is not None else "default (autotune off)"
        line.update({"clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                     "cpu_baseline": cpu_baseline, "loss": float(loss),
                     "interactions_per_s": E / (ms * 1e-3)})
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hm", choices=sorted(WORKLOADS))
    ap.add_argument("--degree", default="powerlaw", choices=["powerlaw", "uniform"])
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--degree-order", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-autotune", action="store_true",
                    help="keep the default SpMM kernel variant instead of timing the candidates on this graph at set-up")
    ap.add_argument("--graph", action="store_true", help="multi-GPU: replay the step from a CUDA graph (opt-in, not yet measured)")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "symm"],
                    help="multi-GPU item-block exchange: NCCL all-reduce (measured default) or the symmetric-memory multimem kernel")
    ap.add_argument("--schedule", default="layer", choices=["layer", "pipelined", "merged"],
                    help="multi-GPU overlap schedule (dist.ShardedLightGCN); 'layer' is the measured default")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
