#!/usr/bin/env bash
# Round-2 GPU call P (ONE GPU): row-sparse first backward layer on both paths (fused_step: batch indices; autograd: rows found by
# lgb_scale_rows_nonzero), resid rows skipped; full GPU suite; bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/p_status.log
run() { local name=$1; shift; timeout "$1" "${@:2}" > $O/$name.log 2>&1; echo "[rc=$?] ${*:2}" >> $O/p_status.log; }
run p_pytest 900 python -m pytest tests -m gpu -q -x --durations=5
run p_probe 400 python tools/spmm_probe.py --variants 23 --chunks 512 --sweep 1 --rowsparse 128
timeout 600 python bench.py --steps 20 --warmup 5 > $O/p_bench_hm.json 2> $O/p_bench_hm.err; echo "[rc=$?] bench hm" >> $O/p_status.log
timeout 300 python bench.py --steps 50 --warmup 5 --workload ml1m --no-cpu-baseline > $O/p_bench_ml1m.json 2> $O/p_bench_ml1m.err
timeout 300 python tools/train_lightgcn.py --workload hm --style fused --sampler device --iters 100 --eval-every 1000 --autotune > $O/p_train_hm_device_sampler_autotune.log 2>&1
timeout 300 python tools/train_lightgcn.py --workload hm --style reference --sampler device --iters 100 --eval-every 1000 --autotune > $O/p_train_hm_reference_style_autotune.log 2>&1
python - >> $O/p_status.log <<'PY'
import json
for f in ("p_bench_hm.json","p_bench_ml1m.json"):
    try:
        j=json.loads([l for l in open("gpurun_out/"+f).read().splitlines() if l.startswith("{")][-1])
        print(f, "ms", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), "loss", j.get("loss"), j["config"]["spmm_variant"].get("forward"), "launches", j["gpu_launches"])
    except Exception as e:
        print(f, "FAILED", e)
for f in ("p_train_hm_device_sampler_autotune.log","p_train_hm_reference_style_autotune.log"):
    try:
        j=json.loads([l for l in open("gpurun_out/"+f).read().splitlines() if l.startswith("{")][-1])
        print(f, {k:j[k] for k in ("iters","iters_per_s","sampler_s","step_s","autotune_s","plan")})
    except Exception as e:
        print(f, "FAILED", e)
PY
cat $O/p_status.log; grep -h "rowsparse" $O/p_probe.log; tail -4 $O/p_pytest.log
