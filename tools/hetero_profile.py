"""Where does the ranking encoder-decoder step spend its time?  torch.profiler over a few training steps of
bench.py's hetero workload: top CPU-side ops (self CPU time), top kernels (CUDA time), launches per step, and the
wall-clock / device-time ratio (host-bound vs device-bound).  NOT a bench number -- profiler overhead is included.

    python tools/hetero_profile.py [--workload hetero_s|hetero_m|hetero_l] [--aggr add|mean] [--steps 10]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="hetero_s")
    ap.add_argument("--aggr", default="add")
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    x, ei, eli, y = bench.hetero_batch(a.workload, dev)
    model, _ = bench.hetero_model(a.aggr)
    model = model.to(dev)
    lossf = torch.nn.BCEWithLogitsLoss()

    def step():
        model.zero_grad(set_to_none=True)
        loss = lossf(model(dict(x), ei, eli), y)
        loss.backward()
        return loss
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    # plain timing first: wall clock with one sync at the end vs CUDA events
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(a.steps):
        step()
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / a.steps * 1e3
    print(f"{a.workload} {a.aggr}: {wall:.3f} ms/step wall, {e0.elapsed_time(e1) / a.steps:.3f} ms/step CUDA events")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
    ka = prof.key_averages()
    cuda_total = sum(getattr(k, "self_device_time_total", getattr(k, "self_cuda_time_total", 0)) for k in ka)
    print(f"device time total {cuda_total / a.steps / 1e3:.3f} ms/step")
    print(ka.table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
    try:
        print(ka.table(sort_by="self_device_time_total", row_limit=25, max_name_column_width=80))
    except Exception:
        print(ka.table(sort_by="self_cuda_time_total", row_limit=25, max_name_column_width=80))


if __name__ == "__main__":
    main()
