#!/usr/bin/env bash
# Round-2 GPU call R (ONE GPU): stage-2 tree fused into the main launch (LGB_SPMM_FUSED_STAGE2) vs the separate tree kernel.
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/r_status.log
run() { local name=$1; shift; timeout "$1" "${@:2}" > $O/$name.log 2>&1; echo "[rc=$?] ${*:2}" >> $O/r_status.log; }
run r_pytest 600 python -m pytest tests/test_gpu_lightgcn.py -q -x -k "fused_stage2 or spmm or golden or small_batch"
LGB_SPMM_STAGE2=fused run r_pytest_fused_default 900 python -m pytest tests -m gpu -q -x
for mode in tree fused; do
  LGB_SPMM_STAGE2=$mode timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/r_bench_hm_$mode.json 2> $O/r_bench_hm_$mode.err; echo "[rc=$?] bench hm $mode" >> $O/r_status.log
  LGB_SPMM_STAGE2=$mode run r_shard_probe8_$mode 300 python tools/shard_probe.py --world 8 --ranks 0 --variants 0,12,20,23 --chunks 256
  LGB_SPMM_STAGE2=$mode run r_probe_$mode 300 python tools/spmm_probe.py --variants 0,23 --chunks 512 --sweep 1 --rowsparse 128
done
python - >> $O/r_status.log <<'PY'
import json
for f in ("r_bench_hm_tree.json","r_bench_hm_fused.json"):
    try:
        j=json.loads([l for l in open("gpurun_out/"+f).read().splitlines() if l.startswith("{")][-1])
        print(f, "ms", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), "loss", j.get("loss"), j["config"]["spmm_variant"].get("forward"), "launches", j["gpu_launches"])
    except Exception as e:
        print(f, "FAILED", e)
PY
cat $O/r_status.log; tail -3 $O/r_pytest.log; tail -3 $O/r_pytest_fused_default.log; for mode in tree fused; do echo $mode; grep -h "^chunk" $O/r_probe_$mode.log; tail -6 $O/r_shard_probe8_$mode.log; done
