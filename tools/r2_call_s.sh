#!/usr/bin/env bash
# Round-2 GPU call S (ONE GPU): final validation of the tree as committed -- full GPU suite, smoke, both bench arms, the other
# workloads, launch list + ncu --set full of the final plan (fused stage 2, column sweep, row-sparse first backward layer).
set -u
mkdir -p gpurun_out
O=gpurun_out
T() { timeout "$@"; echo "[rc=$?] ${*:2}" | cut -c1-200 >> $O/s_status.log; }
: > $O/s_status.log
T 900 python -m pytest tests -m gpu -q -x --durations=5 > $O/s_pytest.log 2>&1
tail -3 $O/s_pytest.log >> $O/s_status.log
T 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/s_smoke.log 2>&1
tail -2 $O/s_smoke.log >> $O/s_status.log
T 600 python bench.py > $O/s_bench_default.json 2> $O/s_bench_default.err
T 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/s_bench_reference.json 2> $O/s_bench_reference.err
eval "$(python tools/pick_winner.py $O/s_bench_default.json)"
echo "winner: v=$FWD_V chunk=$FWD_CHUNK order=$FWD_ORDER hot=$FWD_HOT sweep=$FWD_SWEEP" >> $O/s_status.log
ORD=""; [ "$FWD_ORDER" = "1" ] && ORD="--degree-order"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-autotune $ORD"
export LGB_SPMM_VARIANT=$FWD_V LGB_SPMM_CHUNK=$FWD_CHUNK LGB_SPMM_SWEEP=$FWD_SWEEP
timeout 200 $B > $O/s_plain_winner.json 2> $O/s_plain_winner.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2s_launches_bench_hm.csv \
    $B > $O/s_ncu_launches.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_ -s 6 -c 6 -o $O/r2s_spmm_final \
    $B > $O/s_ncu_full.log 2>&1
echo "[rc=$?] ncu final captures" >> $O/s_status.log
unset LGB_SPMM_VARIANT LGB_SPMM_CHUNK LGB_SPMM_SWEEP
T 300 python bench.py --steps 50 --warmup 5 --workload ml1m > $O/s_bench_ml1m.json 2> $O/s_bench_ml1m.err
for hs in hetero_s hetero_m hetero_l; do
  T 300 python bench.py --workload $hs --steps 100 --warmup 10 --no-cpu-baseline > $O/s_bench_$hs.json 2> $O/s_bench_$hs.err
done
python - >> $O/s_status.log <<'PY'
import json
for f in ("s_bench_default.json","s_bench_reference.json","s_plain_winner.json","s_bench_ml1m.json","s_bench_hetero_s.json","s_bench_hetero_m.json","s_bench_hetero_l.json"):
    try:
        j=json.loads([l for l in open("gpurun_out/"+f).read().splitlines() if l.startswith("{")][-1])
        r=j.get("roofline") or {}
        print(f, "ms", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3) if j.get("e2e") and "ms_per_step" in j["e2e"] else j.get("e2e"), "value", j["value"], "frac", r.get("frac"), "dram_frac", r.get("dram_frac"), "rowsparse", r.get("rowsparse_launches"), "per_step", j.get("per_step_ms") and {k:v for k,v in j["per_step_ms"].items() if k!="note"}, "launches", j.get("gpu_launches"))
    except Exception as e:
        print(f, "FAILED", e)
PY
cat $O/s_status.log
