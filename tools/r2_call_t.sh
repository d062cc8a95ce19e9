#!/usr/bin/env bash
# Round-2 GPU call T (N GPUs): bench.py at N GPUs on the final tree (column sweep in the plan search, row-sparse first backward
# layer, fused stage 2).
set -u
N=${1:-4}
O=gpurun_out
mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus "$N" --steps 20 --warmup 5 > $O/t${N}_bench.json 2> $O/t${N}_bench.err
echo "[rc=$?] bench x$N"
python - $N <<'PY'
import json,sys
N=sys.argv[1]
j=json.loads([l for l in open(f"gpurun_out/t{N}_bench.json").read().splitlines() if l.startswith("{")][-1])
t=j["config"]["spmm_variant"]
print("ms", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), "loss", j.get("loss"), "parity ok", j["shard_parity"]["eager"]["ok"], j["shard_parity"].get("graph",{}).get("ok"),
      {k:{kk:vv for kk,vv in t[k].items() if kk not in ("ms","rejected")} for k in ("users","items")}, "exchange", j["exchange"]["mode"], round(j["exchange"]["mean_ms"],4))
PY
