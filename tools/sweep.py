"""Config 5 of BASELINE.json: embedding-dim x degree-skew sweep of the LightGCN epoch (SpMM + BPR kernels).
Runs bench.py once per point (own timeout each) and prints a table + one JSON line per point.

    python tools/sweep.py [--gpus N] [--dims 32,64,128,256] [--degrees uniform,powerlaw] [--batch 128]
"""
import argparse
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--dims", default="32,64,128,256")
    ap.add_argument("--degrees", default="uniform,powerlaw")
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--timeout", type=int, default=180)
    a = ap.parse_args()
    rows = []
    for degree in a.degrees.split(","):
        for d in [int(x) for x in a.dims.split(",")]:
            base = [sys.executable]
            if a.gpus > 1:
                base += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr", "127.0.0.1",
                         "--master-port", str(29700 + d % 97)]
            cmd = base + [os.path.join(REPO, "bench.py"), "--gpus", str(a.gpus), "--steps", str(a.steps), "--warmup", "3",
                          "--dim", str(d), "--degree", degree, "--batch", str(a.batch), "--no-cpu-baseline"]
            try:
                out = subprocess.run(cmd, capture_output=True, text=True, timeout=a.timeout).stdout.strip().splitlines()
                j = json.loads(out[-1])
                rows.append((degree, d, j["ms_per_step"], j["value"] / 1e9, j["roofline"]["avg_launch_ms"], j["roofline"]["frac"]))
                print(json.dumps({"degree": degree, "d": d, "n_gpus": a.gpus, "ms_per_step": j["ms_per_step"], "value": j["value"],
                                  "spmm_ms": j["roofline"]["avg_launch_ms"], "roofline_frac": j["roofline"]["frac"]}), flush=True)
            except Exception as e:  # a point that fails or times out must not stop the sweep
                print(json.dumps({"degree": degree, "d": d, "error": repr(e)[:200]}), flush=True)
    print(f"{'degree':10s} {'d':>4s} {'ms/epoch':>9s} {'Gtrav/s':>8s} {'spmm ms':>8s} {'frac':>6s}")
    for r in rows:
        print(f"{r[0]:10s} {r[1]:4d} {r[2]:9.3f} {r[3]:8.2f} {r[4]:8.3f} {r[5]:6.2f}")


if __name__ == "__main__":
    main()
