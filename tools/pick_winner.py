"""Print shell assignments naming the SpMM plan bench.py's autotune chose (forward graph): FWD_V FWD_CHUNK FWD_ORDER FWD_HOT FWD_SWEEP.
Used by the GPU call scripts to re-run the SAME configuration under ncu without the autotune launches."""
import json
import sys


def main():
    v, chunk, order, hot, sweep = 0, 1024, 0, 0, 0
    try:
        line = [ln for ln in open(sys.argv[1]).read().splitlines() if ln.startswith("{")][-1]
        ch = json.loads(line)["config"]["spmm_variant"]["forward"]
        v, chunk, order, hot = int(ch["variant"]), int(ch["chunk"]), int(bool(ch["degree_order"])), int(ch.get("hot_rows", 0))
        sweep = int(bool(ch.get("sweep", False)))
    except Exception as exc:       # fall back to the default plan
        print(f"# pick_winner: {exc!r}", file=sys.stderr)
    print(f"FWD_V={v} FWD_CHUNK={chunk} FWD_ORDER={order} FWD_HOT={hot} FWD_SWEEP={sweep}")


if __name__ == "__main__":
    main()
