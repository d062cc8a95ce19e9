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


class CUDA:
    """The CUDA-runtime touch points of this script in one place.  On the GPU box these are the plain torch.cuda calls;
    tests/test_emu_kernels.py swaps the class for wall-clock / CPU stand-ins to dry-run the script's own logic (workload
    set-up, plan-time tuning, JSON assembly) on the emulated kernels -- the product path itself has no such switch."""
    available = staticmethod(lambda: torch.cuda.is_available())
    event = staticmethod(lambda: torch.cuda.Event(enable_timing=True))
    synchronize = staticmethod(lambda: torch.cuda.synchronize())
    empty_cache = staticmethod(lambda: torch.cuda.empty_cache())
    pin = staticmethod(lambda t: t.pin_memory())
    step_timer = None           # None: ShardedLightGCN.autotune_step times with CUDA events
    backend = "nccl"            # torch.distributed backend of the multi-GPU runs
    sharded_ops = None          # None: dist.ShardedLightGCN builds its own CudaOps

    @staticmethod
    def device(local: int) -> torch.device:
        torch.cuda.set_device(local)
        return torch.device("cuda", local)


_ORIG_AFFINITY = None


def unpin_cpus():
    """Give the process all its CPUs back (before the CPU baseline, which uses every host core)."""
    if _ORIG_AFFINITY:
        try:
            os.sched_setaffinity(0, _ORIG_AFFINITY)
        except Exception:
            pass


def pin_to_gpu_cpus(local: int):
    """Run this process on the CPUs that are local to its GPU (sysfs `local_cpulist` of the PCI device, intersected with the
    CPUs the container may use).  The launch-bound steps (ranking model on small batches, ML-1M-sized graphs) are host-dispatch
    bound: on a two-socket box a process that lands on the far socket pays every kernel launch across the socket link.
    Returns a short description for the JSON line; never raises."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        txt = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            if part:
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        global _ORIG_AFFINITY
        _ORIG_AFFINITY = set(allowed)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            return {"gpu_local_cpus": txt, "allowed": len(allowed), "pinned_to": len(use)}
        return {"gpu_local_cpus": txt, "allowed": len(allowed), "pinned_to": None if use else "no local CPU is available to this process"}
    except Exception as exc:
        return {"error": repr(exc)[:120]}


WORKLOADS = {
    # name: (U, I, E)  -- BASELINE.json configs[2] (H&M-shaped) and configs[1] (MovieLens-1M-shaped)
    "hm": (1_371_980, 105_542, 31_788_324),
    "ml1m": (6_040, 3_706, 1_000_209),
}
CPU_SAMPLE_SCALE = 1  # the CPU arms run the SAME graph as the GPU arm (full config: ~3-5 s per epoch on the box's host cores)

# BASELINE.json configs[3]: hetero encoder-decoder ranking step on LinkNeighborLoader-sized batches cut to the H&M shape
# (SURVEY 8d.5): name -> (E_sub per edge type, N_customer, N_article, label edges); widths 84/76 -> hidden 128 -> out 64
HETERO_SIZES = {"hetero_s": (72_000, 3_000, 40_000, 1_100), "hetero_m": (400_000, 16_000, 90_000, 5_800),
                "hetero_l": (3_000_000, 130_000, 105_000, 46_000)}
HET_FC, HET_FA, HET_HID, HET_OUT, HET_LAYERS = 84, 76, 128, 64, 2


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


def spmm_floor_bytes(nnz: int, rows: int, d: int) -> float:
    """Compulsory DRAM bytes of one SpMM launch: CSR arrays once, operand table once, output once."""
    return nnz * 8 + (rows + 1) * 4 + 2 * rows * d * 4


def lookup_traffic(workload, degree, d, n_gpus, plan):
    """ncu-measured DRAM bytes per lgb_spmm call for EXACTLY the configuration that ran (profiles/roofline_traffic.json),
    or None: a capture of another kernel / plan / shard size never stands in."""
    try:
        caps = json.load(open(os.path.join(REPO, "profiles", "roofline_traffic.json")))["captures"]
    except Exception:
        return None
    for c in caps:
        if (c["workload"], c["degree"], c["d"], c["n_gpus"]) == (workload, degree, d, n_gpus) and plan is not None and \
                (c["variant"], c["chunk"], bool(c["degree_order"]), int(c.get("hot_rows", 0)), bool(c.get("sweep", False))) == \
                (plan.get("variant"), plan.get("chunk"), bool(plan.get("degree_order")), int(plan.get("hot_rows", 0)),
                 bool(plan.get("sweep", False))):
            return c
    return None


def epoch_bytes(nnz: int, N: int, d: int, K: int, B: int) -> float:
    bpr = B * 12 * d * 4 + 3 * B * 8 + N * d * 4
    return 2 * K * spmm_bytes(nnz, N, d) + (3 * K + 2) * N * d * 4 + bpr


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region (B200_PROFILING.md).  Read through NVML from a
    thread of this process (pynvml, one query every 25 ms, ~0.1 ms each): an `nvidia-smi -lms` child process was measured to
    slow launch-bound steps down by up to 2x while it polls (its start-up and every poll take driver locks) -- it remains the
    fallback when pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc, self.samples, self.first, self._stop, self.smax = index, [], None, [], 0, False, None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:                                   # CUDA_VISIBLE_DEVICES may renumber: address the device by its PCI bus id
            pr = torch.cuda.get_device_properties(self.index)
            bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            return pynvml, pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def start(self):
        try:
            nv, h = self._nvml_handle()
            self.smax = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}

            def loop():
                while not self._stop:
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), [n for b, n in names.items() if mask & b]))
                    except Exception:
                        pass
                    time.sleep(0.025)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.mode = "nvml"
            return
        except Exception:
            self.mode = "nvidia-smi"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t0 = time.perf_counter()
            while not self.lines and time.perf_counter() - t0 < 5.0:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def mark(self):
        """Samples taken from here on belong to the timed region."""
        self.first = len(self.samples) if getattr(self, "mode", "") == "nvml" else len(self.lines)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if getattr(self, "mode", "") == "nvml":
            time.sleep(0.03)
            self._stop = True
            got = self.samples[max(self.first - 1, 0):]             # the sample that straddles the start counts too
            sm = [c for c, _ in got]
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.smax,
                    "reasons": sorted({r for _, rs in got for r in rs}), "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        lines = self.lines[max(self.first - 1, 0):]
        for ln in lines:
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
                "samples": len(sm), "source": "nvidia-smi"}


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
    sample = (f"the full {args.workload}-shaped graph (U={U}, I={I}, E={E}, nnz={nnz}, {args.degree}; same generator and seed as the GPU arm), "
              f"one full epoch per step, torch CSR (MKL) SpMM + autograd")
    line = base_line(args, value, dt * 1e3, nnz_full=None, sizes=(U, I, E))
    line.update({"impl": "reference", "n_gpus": args.gpus, "dtype": "f32",
                 "cpu_baseline": {"value": value, "unit": "edge-traversals/s", "cores": torch.get_num_threads(),
                                  "kind": "port", "sample": sample},
                 "e2e": {"value": value, "unit": "edge-traversals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)


def base_line(args, value, ms, nnz_full, sizes=None):
    U, I, E = sizes or WORKLOADS[args.workload]     # the sizes that actually ran
    return {
        "metric": "lightgcn_fwd_bwd_bpr_edge_traversals_per_s", "value": value, "unit": "edge-traversals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}_{args.degree}_symmetric", "U": U, "I": I, "E": E, "nnz": 2 * E,
                   "d": args.dim, "K": args.layers, "B": args.batch,
                   "epoch": "K SpMM fwd + layer mean + BPR(B) + K SpMM^T bwd; sampler and Adam excluded",
                   "l2": "working set (tables 378 MB + CSR 1 GB per direction at hm) exceeds the 126 MB L2; no flush",
                   "parallelism": "single GPU" if args.gpus == 1 else f"users range-sharded x{args.gpus}, items replicated, one item-block exchange per layer"},
    }


def time_step(fn, dev, world, dist, reps=10):
    """ms per call, barrier + CUDA events, max over ranks (set-up-time comparisons of step forms)."""
    for _ in range(2):
        fn()
    if world > 1:
        dist.barrier()
    CUDA.synchronize()
    e0, e1 = CUDA.event(), CUDA.event()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    CUDA.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def want_graph(args, eng) -> bool:
    """N > 1: 'auto' captures the step only when the exchange is the package's own kernel (no NCCL call inside the capture)."""
    return args.graph == "on" or (args.graph == "auto" and getattr(eng.ops, "kind", "") == "symm")


def sharded_parity_check(args, dev, rank, world):
    """N > 1: a small fixed problem (skewed degrees, K = 3, d = 64, B = 512) through the SAME engine configuration the bench
    is about to time (schedule, exchange, CUDA-graph replay when requested), every rank's shard against the single-process
    CPU oracle (the bench may call oracle/ as a checker).  Returns the worst relative errors over all ranks."""
    import torch.distributed as dist
    from laplace_gnn_recommendation_b200.dist import ShardedLightGCN
    from oracle import lightgcn_oracle as lo
    gen = torch.Generator().manual_seed(7)
    U, I, E, d, K, B, lam = 4000, 600, 90000, 64, 3, 512, 1e-3
    users = (torch.rand(E, generator=gen) ** 2 * U).long().clamp(max=U - 1)
    items = (torch.rand(E, generator=gen) ** 3 * I).long().clamp(max=I - 1)
    Wu, Wi = torch.randn(U, d, generator=gen) * 0.1, torch.randn(I, d, generator=gen) * 0.1
    pick = torch.randint(0, E, (B,), generator=gen)
    u, p, n = users[pick], items[pick], torch.randint(0, I, (B,), generator=gen)
    ops = CUDA.sharded_ops() if CUDA.sharded_ops else None
    eng = ShardedLightGCN(U, I, d, K, users, items, dev, schedule="chains" if args.schedule == "auto" else args.schedule,
                          exchange=args.exchange, ops=ops, init_tables=(Wu, Wi), max_batch=B)
    ud, pd, nd = u.to(dev), p.to(dev), n.to(dev)
    out = {"exchange": getattr(eng.ops, "kind", "?"), "multicast": bool(getattr(eng.ops, "multicast", False)),
           "mode": getattr(eng.ops, "mode", None), "schedule": eng.schedule}
    forms = {"eager": lambda: eng.fused_step(ud, pd, nd, lam)}
    if want_graph(args, eng):
        try:
            gstep = eng.capture(B, lam)
            forms["graph"] = lambda: gstep(ud, pd, nd)
        except Exception as exc:
            out["graph_error"] = repr(exc)[:200]
    row, col, nn = lo.wiring_symmetric(users, items, U, I)
    rowptr, c, _ = lo.csr_from_coo(row, col, nn, nn)
    o_loss, o_gu, o_gi, o_uf, o_if = lo.train_iteration(Wu, Wi, rowptr, c, K, u, p, n, lam)

    def rel(got, want):
        got, want = got.detach().cpu().double(), want.double()
        return float((got - want).abs().max() / (want.abs().max() + 1e-30))
    for name, fn in forms.items():
        for _ in range(2):                                   # twice: the second pass re-uses every buffer and signal slot
            loss = fn()
        CUDA.synchronize()
        errs = torch.tensor([abs(float(loss) - float(o_loss)) / abs(float(o_loss)),
                             max(rel(eng.E_f_users, o_uf[eng.lo:eng.hi]) if eng.Ug else 0.0, rel(eng.E_f_items, o_if)),
                             max(rel(eng.grad_users, o_gu[eng.lo:eng.hi]) if eng.Ug else 0.0, rel(eng.grad_items, o_gi))],
                            dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        out[name] = {"loss_rel_err": float(errs[0]), "emb_max_rel_err": float(errs[1]), "grad_max_rel_err": float(errs[2]),
                     "ok": bool((errs <= 1e-5).all())}
    out["problem"] = f"U={U} I={I} E={E} d={d} K={K} B={B}, skewed degrees, vs oracle/lightgcn_oracle.train_iteration; errors = max |got-want| / max |want| over all ranks"
    del eng
    CUDA.empty_cache()
    return out


def run_ours(args):
    import torch.distributed as dist

    import laplace_gnn_recommendation_b200 as lg
    from laplace_gnn_recommendation_b200 import _lib
    from laplace_gnn_recommendation_b200.csr import DeviceCSR

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not CUDA.available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    dev = CUDA.device(local)
    affinity = pin_to_gpu_cpus(local) if dev.type == "cuda" else None
    if world > 1:
        dist.init_process_group(CUDA.backend, **({"device_id": dev} if CUDA.backend == "nccl" else {}))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")

    U, I, E = WORKLOADS[args.workload]
    d, K, B, lam = args.dim, args.layers, args.batch, 1e-6
    tuned, shard_parity, graph_report = None, None, None
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
                # the public plan-time call (variant x slice size x row order x slice order, lightgcn.AUTOTUNE_SPACE)
                model.autotune(adj, thorough=True)
                tuned = {"forward": g.autotune_report.get("chosen"), "backward": gt.autotune_report.get("chosen"),
                         "forward_ms": g.autotune_report["ms"], "backward_ms": gt.autotune_report["ms"],
                         "rejected": {**g.autotune_report["rejected"], **gt.autotune_report["rejected"]}}
            except Exception as exc:      # tuning is optional: fall back to the default plan and say so
                g.variant = gt.variant = None
                tuned = {"error": repr(exc)[:300]}
        nnz = g.nnz
        if args.graph == "on":                                   # one CUDA-graph launch per iteration (launch-bound small graphs)
            gstep1 = model.capture_step(adj, B, lam)
            step = lambda: gstep1(ub, pb, nb)                    # noqa: E731
        else:
            step = lambda: model.fused_step(adj, ub, pb, nb, lam)   # noqa: E731
    else:
        from laplace_gnn_recommendation_b200.dist import ShardedLightGCN
        # the SAME embedding tables as the single-GPU model (same generator, same order): the loss printed at N = 1, 2, 4, 8
        # is then one number to rtol 1e-5, which the scaling record can be checked against
        torch.manual_seed(0)
        cpu_model = lg.LightGCN(U, I, d, K)
        tables = (cpu_model.users_emb.weight.detach(), cpu_model.items_emb.weight.detach())
        ops = CUDA.sharded_ops() if CUDA.sharded_ops else None
        eng = ShardedLightGCN(U, I, d, K, users, items, dev, schedule="chains" if args.schedule == "auto" else args.schedule,
                              exchange=args.exchange, ops=ops, init_tables=tables, max_batch=max(B, 128))
        del cpu_model, tables
        shard_parity = sharded_parity_check(args, dev, rank, world)      # small fixed problem through the SAME step form vs the oracle
        if not args.no_autotune:
            try:
                tuned = eng.autotune()                        # per rank; rank 0's choice is reported
            except Exception as exc:
                for gv in eng.graphs():
                    gv.variant = None
                tuned = {"error": repr(exc)[:300]}
            if args.schedule == "auto":
                # schedule (two layer chains on two streams vs one stream): result-checked on every rank, timed max-over-ranks;
                # all ranks run the same candidates in the same order (ShardedLightGCN.autotune_step)
                try:
                    tuned["step_form"] = eng.autotune_step(ub, pb, nb, lam, candidates=("chains", "layer"), timer=CUDA.step_timer)
                except Exception as exc:
                    eng.schedule = "chains"
                    tuned["step_form"] = {"error": repr(exc)[:300]}
        elif args.schedule == "auto":
            eng.schedule = "chains"
        nnz = 2 * E
        eager = lambda: eng.fused_step(ub, pb, nb, lam)          # noqa: E731
        step = eager
        if want_graph(args, eng):                                # whole step (kernels on three streams + exchanges) replayed from one CUDA graph
            try:
                ref_loss = float(eager())
                gstep = eng.capture(B, lam)
                g_loss = float(gstep(ub, pb, nb))
                same = abs(g_loss - ref_loss) <= 1e-5 * abs(ref_loss) + 1e-7
                t_e, t_g = (time_step(f, dev, world, dist) for f in (eager, lambda: gstep(ub, pb, nb)))
                use = same and (args.graph == "on" or t_g < t_e)
                graph_report = {"replay_matches_eager": same, "eager_ms": t_e, "graph_ms": t_g, "used": bool(use)}
                if use:
                    step = lambda: gstep(ub, pb, nb)             # noqa: E731
            except Exception as exc:
                graph_report = {"error": repr(exc)[:300], "used": False}
            if tuned is not None:
                tuned["cuda_graph"] = graph_report
    del users, items
    CUDA.empty_cache()

    # per-launch CUDA-event timing of the dominant kernel (lgb_spmm) inside the timed region
    spmm_events = []
    orig_spmm = DeviceCSR.spmm

    def timed_spmm(self, *a, **k):
        e0, e1 = CUDA.event(), CUDA.event()
        e0.record()
        out = orig_spmm(self, *a, **k)
        e1.record()
        spmm_events.append((e0, e1, self.nnz, self.n_rows, k.get("x_rows") is not None))
        return out

    def sync():
        if world > 1:
            dist.barrier()
        CUDA.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        step()
    sync()
    clocks.mark()
    DeviceCSR.spmm = timed_spmm
    graphed = bool(graph_report and graph_report.get("used")) or (world == 1 and args.graph == "on")
    if world > 1 and hasattr(eng.ops, "exchange_events") and not graphed:
        eng.ops.exchange_events = []           # CUDA events on the comm streams around every exchange
    launches0 = _lib.LAUNCHES
    sync()
    t0, t1 = CUDA.event(), CUDA.event()
    t0.record()
    for _ in range(args.steps):
        loss = step()
    t1.record()
    sync()
    launches = _lib.LAUNCHES - launches0
    clk = clocks.stop() if rank == 0 else None
    if graphed:
        # the timed region replayed a CUDA graph (no per-launch host hooks); per-launch SpMM durations for the
        # roofline and the launch count come from an instrumented kernel-by-kernel pass of the SAME step
        launches0 = _lib.LAUNCHES
        if world > 1 and hasattr(eng.ops, "exchange_events"):
            eng.ops.exchange_events = []
        for _ in range(3):
            (eager() if world > 1 else model.fused_step(adj, ub, pb, nb, lam))
        launches = (_lib.LAUNCHES - launches0) // 3 * args.steps
        sync()
    n_passes = 3 if graphed else args.steps
    exchange = None
    if world > 1 and getattr(eng.ops, "exchange_events", None):
        ev, eng.ops.exchange_events = eng.ops.exchange_events, None
        ev = [e for e in ev if e[2] >= I * d * 4]                    # the item-block exchanges (not the 2*B*d batch rows)
        x_ms = [a.elapsed_time(b) for a, b, _ in ev]
        x_bytes = [n for _, _, n in ev]
        # an all-reduce of M bytes over G ranks moves 2*M*(G-1)/G bytes through each GPU's links (reduce-scatter + all-gather)
        bus = [2.0 * n * (world - 1) / world for n in x_bytes]
        exchange = {"collective": ("own kernel over symmetric memory (lgb_exchange_allreduce_f32: barrier + multimem.ld_reduce / multimem.st + barrier)"
                                   if getattr(eng.ops, "kind", "") == "symm" else "NCCL all-reduce") + " of the replicated item block",
                    "multicast": bool(getattr(eng.ops, "multicast", False)), "mode": getattr(eng.ops, "mode", None),
                    "mode_choice": getattr(eng.ops, "mode_report", None), "per_step": len(ev) // max(n_passes, 1),
                    "mean_ms": statistics.mean(x_ms), "mean_bytes": statistics.mean(x_bytes),
                    "bus_GBps": sum(bus) / 1e9 / (sum(x_ms) * 1e-3), "nvlink_peak_GBps_per_direction": 900.0,
                    "frac_of_nominal": sum(bus) / 1e9 / (sum(x_ms) * 1e-3) / 900.0,
                    "share_of_step_if_exposed": sum(x_ms) / n_passes / (float(t0.elapsed_time(t1)) / max(args.steps, 1)),
                    "timed_in": "separate kernel-by-kernel pass (timed region replays a CUDA graph)" if graphed else "timed region",
                    "measured_reference_GBps": 770.0,
                    "note": "timed on the high-priority comm streams (the time includes waiting for the slowest rank at the entry barrier); "
                            "it overlaps with the other chain's SpMM launches"}
    DeviceCSR.spmm = orig_spmm
    ms = torch.tensor([t0.elapsed_time(t1) / max(args.steps, 1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    value = 2 * K * nnz / (ms * 1e-3)

    # the roofline is the DENSE lgb_spmm launch's; the row-sparse first backward layer (lgb_spmm_rowsparse) skips most of the
    # gathers the row-gather model counts, so it is timed and reported next to it, not folded into the byte rate
    sparse_ms = [a.elapsed_time(b) for a, b, _, _, sp in spmm_events if sp]
    spmm_events = [ev[:4] for ev in spmm_events if not ev[4]]
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
    avg_ms = statistics.mean(spmm_ms) if spmm_ms else None
    plan = None
    if world == 1:        # the plan the forward launches ran with (autotune's choice, or the env / flag defaults)
        plan = {"variant": g.variant if g.variant is not None else int(os.environ.get("LGB_SPMM_VARIANT", "0")),
                "chunk": g.chunk, "degree_order": g.row_order is not None, "hot_rows": int(getattr(g, "n_hot", 0)),
                "sweep": bool(getattr(g, "sweep", False))}
    cap = lookup_traffic(args.workload, args.degree, d, world, plan)
    traffic = cap["dram_bytes_per_call"] if cap else None
    floor = spmm_floor_bytes(spmm_events[0][2], spmm_events[0][3], d) if spmm_events else None
    roofline = {"bound": "hbm", "kernel": "lgb_spmm (lgb::spmm_subwarp_kernel / spmm_rows_kernel + spmm_long_reduce_kernel, plan per config.spmm_variant)",
                "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "frac_note": "algorithmic row-gather model (SURVEY 8d): every gathered row counted as DRAM traffic, so > 1 when rows are served "
                             "by L1/L2; dram_frac is the physical figure",
                # physical DRAM bandwidth of the dominant kernel: ncu dram bytes of the SAME plan / live CUDA-event launch time / peak
                "dram_frac": (traffic / 1e9) / (avg_ms * 1e-3) / peak if (traffic and avg_ms) else None,
                "traffic_source": cap["source"] if cap else "no ncu capture of this exact configuration (workload, degree, d, n_gpus, plan)",
                "l2_hit_pct": cap.get("lts_hit_pct") if cap else None, "l1_hit_pct": cap.get("l1_hit_pct") if cap else None,
                # what bounds the launch once the rows come from L2: bytes that cross L2 -> SM (ncu l1tex__m_xbar2l1tex_read_bytes of
                # the same plan) / live launch time; ncu rates it against 2 sectors / clk / L2 slice, B300_MICROARCH.md measures
                # ~6 300 B / clk (~12.4 TB/s at 1.965 GHz) as what the L2 -> SM path sustains
                "l2_to_sm": ({"bytes_per_launch": cap["l2_to_sm_bytes_per_call"],
                              "GBps": cap["l2_to_sm_bytes_per_call"] / 1e9 / (avg_ms * 1e-3),
                              "frac_of_sustained_12400_GBps": cap["l2_to_sm_bytes_per_call"] / 1e9 / (avg_ms * 1e-3) / 12400.0,
                              "ncu_pct_of_l2_read_peak": cap.get("l2_read_pct_of_ncu_peak")}
                             if (cap and cap.get("l2_to_sm_bytes_per_call") and avg_ms) else None),
                "compulsory_floor_bytes_per_launch": floor,
                "traffic_over_floor": traffic / floor if (traffic and floor) else None,
                "plan": plan,
                "algorithmic_bytes_per_launch": spmm_alg[0] if spmm_alg else None,
                "avg_launch_ms": avg_ms,
                "launches_timed": len(spmm_ms),
                "timed_in": "separate kernel-by-kernel pass (timed region replays a CUDA graph)" if graphed else "timed region",
                "spmm_share_of_step": ((sum(spmm_ms) + sum(sparse_ms)) / ((3 if graphed else args.steps) * ms)) if spmm_ms else None,
                "rowsparse_launches": {"launches_timed": len(sparse_ms), "avg_launch_ms": statistics.mean(sparse_ms) if sparse_ms else None,
                                       "note": "first backward layer: dE_f is non-zero on the <= 3B batch rows (lgb_spmm_rowsparse); not part of achieved / frac"},
                "epoch_algorithmic_gb": epoch_bytes(nnz, U + I, d, K, B) / 1e9,
                "epoch_frac_of_peak": epoch_bytes(nnz, U + I, d, K, B) / 1e9 / (ms * 1e-3) / peak}

    # ---- e2e: the reference-style call sequence through the public API, host buffers in the timed region
    e2e = None
    if world == 1:
        hu, hp, hn = (CUDA.pin(t.cpu()) for t in (ub, pb, nb))

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
        CUDA.synchronize()
        e_ms = (time.perf_counter() - w0) * 1e3 / max(args.steps, 1)
        e2e = {"value": 2 * K * nnz / (e_ms * 1e-3), "unit": "edge-traversals/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": 3 * B * 8, "d2h_bytes_per_step": 4,
               "path": "LightGCN.forward(SparseTensor) -> 6 gathers -> bpr_loss -> loss.backward() -> loss.item()"}
    else:
        hu, hp, hn = (CUDA.pin(t.cpu()) for t in (ub, pb, nb))

        def api_step():
            if graphed:
                return gstep(hu, hp, hn).item()
            u_ = hu.to(dev, non_blocking=True); p_ = hp.to(dev, non_blocking=True); n_ = hn.to(dev, non_blocking=True)
            return eng.fused_step(u_, p_, n_, lam).item()
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
               "path": "ShardedLightGCN.fused_step(host-pinned batch) -> loss.item()" + (" [CUDA-graph replay]" if graphed else "")}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        unpin_cpus()
        torch.set_num_threads(os.cpu_count() or 1)
        Us, Is, Es = (x // CPU_SAMPLE_SCALE for x in WORKLOADS[args.workload])
        cstep, cnnz = cpu_epoch_runner(Us, Is, Es, d, K, B, args.degree, 1234)
        cstep()
        c0 = time.perf_counter()
        reps = 0
        while reps < 2 or (time.perf_counter() - c0 < 10 and reps < 20):
            cstep(); reps += 1
        cdt = (time.perf_counter() - c0) / reps
        cpu_baseline = {"value": 2 * K * cnnz / cdt, "unit": "edge-traversals/s", "cores": torch.get_num_threads(),
                        "kind": "port", "ms_per_step": cdt * 1e3,
                        "sample": f"the full {args.workload}-shaped graph (U={Us}, I={Is}, E={Es}, nnz={cnnz}, {args.degree}; same generator "
                                  f"and seed as the GPU arm), {reps} full epochs after 1 warm-up; oracle port, torch CSR (MKL) SpMM + autograd"}

    if rank == 0:
        line = base_line(args, value, ms, nnz)
        line["config"]["cpu_affinity"] = affinity
        line["config"]["spmm_variant"] = tuned if tuned is not None else "default (autotune off)"
        if exchange is not None:
            line["exchange"] = exchange
        if shard_parity is not None:
            line["shard_parity"] = shard_parity
        line.update({"clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                     "cpu_baseline": cpu_baseline, "loss": float(loss),
                     "interactions_per_s": E / (ms * 1e-3)})
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# hetero encoder-decoder ranking step (BASELINE.json configs[3]); python bench.py --workload hetero_m
def hetero_batch(size, dev, seed=0):
    from laplace_gnn_recommendation_b200 import hetero
    E, Nc, Na, L = HETERO_SIZES[size]
    gen = torch.Generator().manual_seed(seed)
    x = {"customer": torch.randn(Nc, HET_FC, generator=gen), "article": torch.randn(Na, HET_FA, generator=gen)}
    e = torch.stack([torch.randint(0, Nc, (E,), generator=gen), torch.randint(0, Na, (E,), generator=gen)])
    ei = {hetero.EDGE_KEY: e, hetero.REV_EDGE_KEY: e.flip(0).contiguous()}
    eli = torch.stack([torch.randint(0, Nc, (L,), generator=gen), torch.randint(0, Na, (L,), generator=gen)])
    y = (torch.rand(L, generator=gen) > 0.75).float()
    to = (lambda t: t.to(dev)) if dev is not None else (lambda t: t)
    return {k: to(v) for k, v in x.items()}, {k: to(v) for k, v in ei.items()}, to(eli), to(y)


def seg_bytes(E, F, n_dst):
    """Algorithmic bytes of one neighbour aggregation: E*(4 + F*4) + (N_dst+1)*4 + N_dst*F*4."""
    return E * (4 + F * 4) + (n_dst + 1) * 4 + n_dst * F * 4


def hetero_model(aggr, project_first=False):
    import laplace_gnn_recommendation_b200 as lg
    from laplace_gnn_recommendation_b200 import hetero
    metadata = (["customer", "article"], [hetero.EDGE_KEY, hetero.REV_EDGE_KEY])
    torch.manual_seed(0)
    model = lg.Encoder_Decoder_Model(
        encoder_layers=lg.get_SAGEConv_layers(HET_LAYERS, HET_HID, HET_OUT, aggr),
        decoder_layers=lg.get_linear_layers(2, 2 * HET_OUT, HET_HID, 1), feature_info={}, metadata=metadata, embedding=False,
        heterogeneous_prop_agg_type="sum", batch_normalize=True, p_dropout_edges=None, p_dropout_features=None)
    if project_first:
        for m in model.modules():
            if isinstance(m, hetero.SAGEConv):
                m.project_first = True
    return model, metadata


def hetero_cpu_step_runner(size, aggr, state_dict, metadata):
    """The CPU arm: the oracle's training step (hetero SAGE encoder, batch norm, concat-MLP decoder, BCE, autograd backward)."""
    from oracle import hetero_oracle as ho
    x, ei, eli, y = hetero_batch(size, None)
    sd = {k: v.detach().cpu().clone().requires_grad_(v.dtype.is_floating_point) for k, v in state_dict.items()}
    layers = [{et: dict(w_l=sd[f"encoder.layers.{li}.{'__'.join(et)}.lin_l.weight"], b_l=sd[f"encoder.layers.{li}.{'__'.join(et)}.lin_l.bias"],
                        w_r=sd[f"encoder.layers.{li}.{'__'.join(et)}.lin_r.weight"]) for et in metadata[1]} for li in range(HET_LAYERS)]
    lin = [(sd[f"decoder.layers.{i}.weight"], sd[f"decoder.layers.{i}.bias"]) for i in range(2)]

    def step():
        z = ho.hetero_encoder(x, ei, layers, aggr, "sum", metadata[1])
        zu = ho.batch_norm_train(z["customer"], sd["encoder_layer_norm_customer.weight"], sd["encoder_layer_norm_customer.bias"])
        zi = ho.batch_norm_train(z["article"], sd["encoder_layer_norm_article.weight"], sd["encoder_layer_norm_article.bias"])
        loss = ho.bce_with_logits(ho.edge_decoder_mlp(zu, zi, eli, lin), y)
        loss.backward()
        return loss
    return step


def hetero_line(args, value, ms):
    E, Nc, Na, L = HETERO_SIZES[args.workload]
    return {"metric": "hetero_sage_ranking_step_edge_traversals_per_s", "value": value, "unit": "edge-traversals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}_{args.hetero_aggr}", "E_sub_per_edge_type": E, "N_customer": Nc, "N_article": Na,
                       "label_edges": L, "widths": [HET_FC, HET_FA, HET_HID, HET_OUT], "sage_layers": HET_LAYERS,
                       "step": "encoder (2 edge types x 2 SAGE layers) + batch norm + concat-MLP decoder + BCE, forward + backward",
                       "traversals_per_step": "2 edge types x E_sub x layers x (fwd + bwd)",
                       "l2": "L: features + messages exceed L2; S/M are L2-resident (stated, not flushed: the reference's own batch sizes)",
                       "project_first": bool(args.project_first),
                       "parallelism": "single GPU" if args.gpus == 1 else f"independent batches x{args.gpus}, gradient all-reduce of the dense weights"}}


def run_hetero(args):
    import torch.distributed as dist
    from laplace_gnn_recommendation_b200 import _lib
    from laplace_gnn_recommendation_b200.csr import DeviceCSR
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not CUDA.available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    dev = CUDA.device(local)
    affinity = pin_to_gpu_cpus(local) if dev.type == "cuda" else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    E, Nc, Na, L = HETERO_SIZES[args.workload]
    x, ei, eli, y = hetero_batch(args.workload, dev, seed=rank)
    model, metadata = hetero_model(args.hetero_aggr, args.project_first)
    model = model.to(dev)
    lossf = torch.nn.BCEWithLogitsLoss()
    params = None

    def step(xb=x, eib=ei, elib=eli, yb=y):
        model.zero_grad(set_to_none=True)
        loss = lossf(model(dict(xb), eib, elib), yb)
        loss.backward()
        if world > 1:                      # data-parallel replicas: one all-reduce of the (small, dense) weight gradients
            flat = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
            dist.all_reduce(flat)
        return loss

    def sync():
        if world > 1:
            dist.barrier()
        CUDA.synchronize()
    events = []
    orig = DeviceCSR.spmm

    def timed(self, *a, **k):
        e0, e1 = CUDA.event(), CUDA.event()
        e0.record(); out = orig(self, *a, **k); e1.record()
        events.append((e0, e1, self.nnz, self.n_rows, a[0].shape[1]))
        return out
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        step()
    sync()
    clocks.mark()
    DeviceCSR.spmm = timed
    launches0 = _lib.LAUNCHES
    sync()
    t0, t1 = CUDA.event(), CUDA.event()
    marks = [CUDA.event() for _ in range(args.steps)]          # one event per step: median / max next to the contract's mean
    import gc
    # A full (generation-2) collection walks every object torch and the bench have alive -- tens of ms for a 2 ms host-bound
    # step.  gc.freeze() moves what exists now out of the collector's reach (the production recipe for long-running servers);
    # the collector itself stays ON, so garbage the steps create is still collected.  LGB_BENCH_GC=default|disable to compare.
    gc_mode = os.environ.get("LGB_BENCH_GC", "default")
    gc.collect()
    if gc_mode == "freeze":
        gc.freeze()
    elif gc_mode == "disable":
        gc.disable()
    t0.record()
    for i in range(args.steps):
        loss = step()
        marks[i].record()
    t1 = marks[-1] if marks else t1
    if not marks:
        t1.record()
    sync()
    gc.enable()
    if gc_mode == "freeze":
        gc.unfreeze()
    DeviceCSR.spmm = orig
    launches = _lib.LAUNCHES - launches0
    per_step = [a.elapsed_time(b) for a, b in zip([t0] + marks[:-1], marks)]
    clk = clocks.stop() if rank == 0 else None
    ms = torch.tensor([t0.elapsed_time(t1) / max(args.steps, 1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    trav = 2 * E * HET_LAYERS * 2
    value = world * trav / (ms * 1e-3)
    agg_ms = [a.elapsed_time(b) for a, b, *_ in events]
    agg_alg = [seg_bytes(z, F_, r) for _, _, z, r, F_ in events]
    mp = {}
    try:
        mp = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(mp.get("hbm_gbs", 6650.0))
    achieved = (sum(agg_alg) / 1e9) / (sum(agg_ms) * 1e-3) if agg_ms else None
    roofline = {"bound": "hbm", "kernel": "lgb_spmm with val == NULL (per-edge-type neighbour sum/mean, forward and transposed backward)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": None,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in mp else "fallback 6650 GB/s (B200_PROFILING.md)",
                "algorithmic_bytes_per_step": sum(agg_alg) / max(args.steps, 1), "launches_timed": len(agg_ms),
                "avg_launch_ms": statistics.mean(agg_ms) if agg_ms else None,
                "aggregation_share_of_step": sum(agg_ms) / (args.steps * ms) if agg_ms else None}

    # ---- e2e: the batch arrives in pinned host memory (what a DataLoader hands over), copied inside the timed region
    hx, hei, heli, hy = hetero_batch(args.workload, None, seed=rank)
    pin = CUDA.pin
    hx, hei, heli, hy = {k: pin(v) for k, v in hx.items()}, {k: pin(v) for k, v in hei.items()}, pin(heli), pin(hy)
    h2d = sum(t.numel() * t.element_size() for t in list(hx.values()) + list(hei.values()) + [heli, hy])

    def api_step():
        xb = {k: v.to(dev, non_blocking=True) for k, v in hx.items()}
        eib = {k: v.to(dev, non_blocking=True) for k, v in hei.items()}
        return step(xb, eib, heli.to(dev, non_blocking=True), hy.to(dev, non_blocking=True)).item()
    for _ in range(3):
        api_step()
    sync()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        api_step()
    sync()
    e_ms = torch.tensor([(time.perf_counter() - w0) * 1e3 / max(args.steps, 1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
    e2e = {"value": world * trav / (float(e_ms) * 1e-3), "unit": "edge-traversals/s", "ms_per_step": float(e_ms),
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
           "path": "Encoder_Decoder_Model.forward(x_dict, edge_index_dict, edge_label_index) -> BCEWithLogitsLoss -> backward -> loss.item()"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        unpin_cpus()
        torch.set_num_threads(os.cpu_count() or 1)
        csize = "hetero_s"                                   # bounded sample: the S batch whatever size the GPU arm ran
        cstep = hetero_cpu_step_runner(csize, args.hetero_aggr, model.state_dict(), metadata)
        cstep()
        c0, reps = time.perf_counter(), 0
        while reps < 3 or (time.perf_counter() - c0 < 10 and reps < 50):
            cstep(); reps += 1
        cdt = (time.perf_counter() - c0) / reps
        ctrav = 2 * HETERO_SIZES[csize][0] * HET_LAYERS * 2
        cpu_baseline = {"value": ctrav / cdt, "unit": "edge-traversals/s", "cores": torch.get_num_threads(), "kind": "port",
                        "ms_per_step": cdt * 1e3, "sample": f"{csize} batch (E_sub={HETERO_SIZES[csize][0]}), {reps} full training steps after 1 "
                                                              "warm-up; oracle port (index_add scatter, torch Linear, autograd)"}
    if rank == 0:
        line = hetero_line(args, value, ms)
        line["config"]["cpu_affinity"] = affinity
        line["config"]["gc"] = gc_mode
        line.update({"clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
                     "loss": float(loss.detach()),
                     "per_step_ms": {"median": statistics.median(per_step), "min": min(per_step), "max": max(per_step),
                                     "steps_over_5x_median": sum(1 for x in per_step if x > 5 * statistics.median(per_step)),
                                     "note": "this step is bound by host dispatch on the S / M batches: the mean (ms_per_step, the "
                                             "contract's number) moves with the host, the median is the steadier figure"} if per_step else None})
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_hetero_reference(args):
    """--impl reference for the hetero workloads: the oracle's training step on the S batch, all host cores."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    from laplace_gnn_recommendation_b200 import hetero
    metadata = (["customer", "article"], [hetero.EDGE_KEY, hetero.REV_EDGE_KEY])
    sd = materialized_state_dict()
    step = hetero_cpu_step_runner("hetero_s", args.hetero_aggr, sd, metadata)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = 2 * HETERO_SIZES["hetero_s"][0] * HET_LAYERS * 2 / dt
    line = hetero_line(args, value, dt * 1e3)
    line.update({"impl": "reference",
                 "cpu_baseline": {"value": value, "unit": "edge-traversals/s", "cores": torch.get_num_threads(), "kind": "port",
                                  "sample": "hetero_s batch, full training step per step; oracle port"},
                 "e2e": {"value": value, "unit": "edge-traversals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)


def materialized_state_dict():
    """Weights of the lazily-sized hetero model without a device: plain tensors of the right shapes, same initialisers."""
    from laplace_gnn_recommendation_b200 import hetero
    torch.manual_seed(0)
    sd = {}
    widths = {"customer": HET_FC, "article": HET_FA}
    for li in range(HET_LAYERS):
        out = HET_HID if li < HET_LAYERS - 1 else HET_OUT
        for et in (hetero.EDGE_KEY, hetero.REV_EDGE_KEY):
            name = f"encoder.layers.{li}.{'__'.join(et)}"
            lin_l, lin_r = torch.nn.Linear(widths[et[0]], out), torch.nn.Linear(widths[et[2]], out, bias=False)
            sd[f"{name}.lin_l.weight"], sd[f"{name}.lin_l.bias"], sd[f"{name}.lin_r.weight"] = lin_l.weight, lin_l.bias, lin_r.weight
        widths = {"customer": out, "article": out}
    for t in ("customer", "article"):
        sd[f"encoder_layer_norm_{t}.weight"], sd[f"encoder_layer_norm_{t}.bias"] = torch.ones(HET_OUT), torch.zeros(HET_OUT)
    dec = [torch.nn.Linear(2 * HET_OUT, HET_HID), torch.nn.Linear(HET_HID, 1)]
    for i, l in enumerate(dec):
        sd[f"decoder.layers.{i}.weight"], sd[f"decoder.layers.{i}.bias"] = l.weight, l.bias
    return sd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hm", choices=sorted(WORKLOADS) + sorted(HETERO_SIZES),
                    help="hm / ml1m: LightGCN epoch (BASELINE.json configs[2] / [1]); hetero_s|m|l: ranking encoder-decoder step (configs[3])")
    ap.add_argument("--hetero-aggr", default="add", choices=["add", "mean"], help="SAGEConv aggregation of the hetero workloads")
    ap.add_argument("--project-first", action="store_true", help="hetero workloads: SAGEConv.project_first (opt-in evaluation order)")
    ap.add_argument("--degree", default="powerlaw", choices=["powerlaw", "uniform"])
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--degree-order", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-autotune", action="store_true",
                    help="keep the default SpMM kernel variant instead of timing the candidates on this graph at set-up")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the step from one CUDA graph: N=1 only when 'on' (launch-bound small graphs); N>1 'auto' captures the step, "
                         "requires the replay to reproduce the eager loss and keeps whichever is faster")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "symm"],
                    help="multi-GPU item-block exchange: own kernel over symmetric memory (multimem / peer), NCCL all-reduce, or auto = symm when it sets up")
    ap.add_argument("--schedule", default="auto", choices=["auto", "chains", "layer"],
                    help="multi-GPU schedule (dist.ShardedLightGCN): the two layer chains on two streams, one stream, or auto = time both at set-up")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload in HETERO_SIZES:
        (run_hetero_reference if args.impl == "reference" else run_hetero)(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
