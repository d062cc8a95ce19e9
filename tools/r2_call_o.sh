#!/usr/bin/env bash
# Round-2 GPU call O (ONE GPU): row-sparse first backward layer (lgb_spmm_rowsparse) -- parity, probe, bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/o_status.log
run() { local name=$1; shift; timeout "$1" "${@:2}" > $O/$name.log 2>&1; echo "[rc=$?] ${*:2}" >> $O/o_status.log; }
run o_pytest 400 python -m pytest tests/test_gpu_lightgcn.py -q -x -k "rowsparse or golden or sweep or oracle"
run o_probe 400 python tools/spmm_probe.py --variants 23 --chunks 512 --sweep 1 --rowsparse 128
run o_probe_b4096 400 python tools/spmm_probe.py --variants 23 --chunks 512 --sweep 1 --rowsparse 4096
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/o_bench_hm.json 2> $O/o_bench_hm.err; echo "[rc=$?] bench hm" >> $O/o_status.log
LGB_ROWSPARSE_BACKWARD=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/o_bench_hm_dense_backward.json 2> $O/o_bench_hm_dense_backward.err; echo "[rc=$?] bench hm dense backward" >> $O/o_status.log
timeout 300 python bench.py --steps 50 --warmup 5 --workload ml1m --no-cpu-baseline > $O/o_bench_ml1m.json 2> $O/o_bench_ml1m.err
python - >> $O/o_status.log <<'PY'
import json
for f in ("o_bench_hm.json","o_bench_hm_dense_backward.json","o_bench_ml1m.json"):
    try:
        j=json.loads([l for l in open("gpurun_out/"+f).read().splitlines() if l.startswith("{")][-1])
        print(f, "ms", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), "loss", j.get("loss"), j["config"]["spmm_variant"].get("forward"))
    except Exception as e:
        print(f, "FAILED", e)
PY
cat $O/o_status.log; grep -h "rowsparse" $O/o_probe.log $O/o_probe_b4096.log; tail -3 $O/o_pytest.log
