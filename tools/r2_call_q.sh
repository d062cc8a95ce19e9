#!/usr/bin/env bash
# Round-2 GPU call Q (N GPUs): sharded engine with the column sweep in its plan search and the row-sparse first backward layer:
# hardware parity (dist_gpu_check) + bench.py at N GPUs, rowsparse on and off.
set -u
N=${1:-2}
O=gpurun_out
mkdir -p $O
S=$O/q${N}_status.log
: > $S
TR() { local t=$1; shift; timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; local rc=$?; echo "[rc=$rc] torchrun x$N $*" | cut -c1-200 >> $S; return $rc; }
DIST_CHECK_ROWSPARSE=1 DIST_CHECK_MODE=chains,symm,graph TR 300 tests/dist_gpu_check.py > $O/q${N}_dist_check.log 2>&1
DIST_CHECK_ROWSPARSE=1 DIST_CHECK_MODE=layer,nccl TR 300 tests/dist_gpu_check.py >> $O/q${N}_dist_check.log 2>&1
DIST_CHECK_MODE=chains,auto TR 300 tests/dist_gpu_check.py >> $O/q${N}_dist_check.log 2>&1
grep -h "DIST_OK\|DIST_FAIL\|Error" $O/q${N}_dist_check.log | head -8 >> $S
TR 300 bench.py --gpus "$N" --steps 20 --warmup 5 > $O/q${N}_bench.json 2> $O/q${N}_bench.err
LGB_ROWSPARSE_BACKWARD=0 TR 300 bench.py --gpus "$N" --steps 20 --warmup 5 > $O/q${N}_bench_dense_backward.json 2> $O/q${N}_bench_dense_backward.err
python - $N >> $S <<'PY'
import json,sys
N=sys.argv[1]
for f in (f"q{N}_bench.json", f"q{N}_bench_dense_backward.json"):
    try:
        j=json.loads([l for l in open("gpurun_out/"+f).read().splitlines() if l.startswith("{")][-1])
        print(f, "ms", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), "loss", j.get("loss"), "parity", j.get("shard_parity"), "tuned", json.dumps(j["config"].get("spmm_variant"))[:400])
    except Exception as e:
        print(f, "FAILED", e)
PY
cat $S
