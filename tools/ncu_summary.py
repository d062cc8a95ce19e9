"""Turn an .ncu-rep (brought back in gpurun_out/) into the short text summary committed under profiles/.

    python tools/ncu_summary.py gpurun_out/spmm_r2.ncu-rep > profiles/r2_spmm_ncu_full_summary.txt
    python tools/ncu_summary.py --json gpurun_out/spmm_r2.ncu-rep      # also prints {"traffic": bytes per launch} for roofline_traffic.json

Runs `ncu -i <rep> --page raw --csv` here (no GPU needed) and keeps the metrics the roofline discussion uses.
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__shared_mem_per_block_static",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum",
]


def to_bytes(value: str, unit: str) -> float:
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return float(value.replace(",", "")) * scale


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    want_json = "--json" in sys.argv
    rep = args[0]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    if out.returncode != 0:
        sys.exit(out.stderr[-2000:])
    rows = list(csv.reader(io.StringIO(out.stdout)))
    header, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(header)}
    traffic = []
    for r in data:
        name = r[col["Kernel Name"]]
        grid = r[col["Grid Size"]] if "Grid Size" in col else ""
        print(f"== {name}  grid {grid}")
        for m in KEEP:
            if m in col and r[col[m]] != "":
                print(f"   {m} = {r[col[m]]} {units[col[m]]}")
        if "dram__bytes_read.sum" in col:
            traffic.append(to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
                           + to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]]))
    if want_json and traffic:
        print(json.dumps({"traffic_bytes_per_launch_mean": sum(traffic) / len(traffic), "launches": len(traffic)}))


if __name__ == "__main__":
    main()
