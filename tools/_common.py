"""Shared helpers of the measurement tools (tools/*.py are NOT product code).

LGB_TOOLS_DRYRUN=1 runs a tool end to end WITHOUT a GPU: tensors live on the CPU, the library is the CPU emulation of the
kernels (tests/emu/) and "CUDA events" are wall-clock stamps.  The numbers mean nothing then -- the point is to catch typos
and shape errors in a probe before it spends GPU-minutes.  Workload sizes are scaled down by the tool when DRYRUN is set.
"""
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

DRYRUN = os.environ.get("LGB_TOOLS_DRYRUN") == "1"
_ctx = None
if DRYRUN:
    from tests.emu.harness import emulated
    _ctx = emulated()
    _ctx.__enter__()          # stays active for the life of the process


def device(index: int = 0) -> torch.device:
    return torch.device("cpu") if DRYRUN else torch.device("cuda", index)


class _WallEvent:
    def __init__(self):
        self.t = 0.0

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other) -> float:
        return (other.t - self.t) * 1e3


def Event():
    return _WallEvent() if DRYRUN else torch.cuda.Event(enable_timing=True)


def sync():
    if not DRYRUN:
        torch.cuda.synchronize()


def empty_cache():
    if not DRYRUN:
        torch.cuda.empty_cache()


def timeit(fn, reps=20, warmup=5):
    """ms per call, CUDA events on the current stream around `reps` back-to-back calls."""
    if DRYRUN:
        reps, warmup = 1, 1
    for _ in range(warmup):
        fn()
    sync()
    e0, e1 = Event(), Event()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    sync()
    return e0.elapsed_time(e1) / reps


def make_ops(dev):
    """dist.CudaOps on the GPU; in a dry run the same class on the emulated library without its CUDA streams."""
    if DRYRUN:
        from tests.test_dist_gloo import make_emu_ops
        return make_emu_ops()
    from laplace_gnn_recommendation_b200.dist import CudaOps
    return CudaOps(dev)
