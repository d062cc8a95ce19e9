#!/usr/bin/env bash
# Round-2 GPU call N (ONE GPU): full GPU test suite on the tree with the column-sweep plan, bench, launch list + ncu --set full of
# the plan the autotune now chooses (variant 23, slices of 512, degree order, column sweep).
set -u
mkdir -p gpurun_out
O=gpurun_out
T() { timeout "$@"; echo "[rc=$?] ${*:2}" | cut -c1-200 >> $O/n_status.log; }
: > $O/n_status.log
T 900 python -m pytest tests -m gpu -q -x --durations=5 > $O/n_pytest.log 2>&1
tail -4 $O/n_pytest.log >> $O/n_status.log
T 500 python bench.py --steps 20 --warmup 5 > $O/n_bench_hm.json 2> $O/n_bench_hm.err
eval "$(python tools/pick_winner.py $O/n_bench_hm.json)"
echo "winner: v=$FWD_V chunk=$FWD_CHUNK order=$FWD_ORDER hot=$FWD_HOT sweep=$FWD_SWEEP" >> $O/n_status.log
ORD=""; [ "$FWD_ORDER" = "1" ] && ORD="--degree-order"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-autotune $ORD"
export LGB_SPMM_VARIANT=$FWD_V LGB_SPMM_CHUNK=$FWD_CHUNK LGB_SPMM_SWEEP=$FWD_SWEEP
timeout 200 $B > $O/n_plain_winner.json 2> $O/n_plain_winner.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2n_launches_bench_hm.csv \
    $B > $O/n_ncu_launches.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_ -s 8 -c 4 -o $O/r2n_spmm_winner \
    $B > $O/n_ncu_full.log 2>&1
echo "[rc=$?] ncu winner captures" >> $O/n_status.log
unset LGB_SPMM_VARIANT LGB_SPMM_CHUNK LGB_SPMM_SWEEP
T 300 python bench.py --steps 50 --warmup 5 --workload ml1m --no-cpu-baseline > $O/n_bench_ml1m.json 2> $O/n_bench_ml1m.err
T 300 python bench.py --steps 20 --warmup 5 --degree uniform --no-cpu-baseline > $O/n_bench_hm_uniform.json 2> $O/n_bench_hm_uniform.err
T 300 python tools/train_lightgcn.py --workload hm --style fused --sampler device --iters 50 --eval-every 1000 > $O/n_train_hm_device_sampler.log 2>&1
cat $O/n_status.log
