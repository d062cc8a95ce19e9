#!/usr/bin/env bash
# Round-2 GPU call F (ONE GPU): stage-2 (long-row reduce) with 1024-thread CTAs + slice size as a wider plan dimension.
set -u
mkdir -p gpurun_out
O=gpurun_out
T() { timeout "$@"; echo "[rc=$?] ${*:2}" | cut -c1-200 >> $O/f_status.log; }
: > $O/f_status.log
T 900 python -m pytest tests -m gpu -q -x --durations=5 > $O/f_pytest.log 2>&1
tail -8 $O/f_pytest.log >> $O/f_status.log
T 500 python bench.py --steps 20 --warmup 5 > $O/f_bench_hm.json 2> $O/f_bench_hm.err
eval "$(python tools/pick_winner.py $O/f_bench_hm.json)"
echo "winner: v=$FWD_V chunk=$FWD_CHUNK order=$FWD_ORDER hot=$FWD_HOT" >> $O/f_status.log
ORD=""; [ "$FWD_ORDER" = "1" ] && ORD="--degree-order"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-autotune $ORD"
export LGB_SPMM_VARIANT=$FWD_V LGB_SPMM_CHUNK=$FWD_CHUNK
timeout 200 $B > $O/f_plain_winner.json 2> $O/f_plain_winner.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2f_launches_bench_hm.csv \
    $B > $O/f_ncu_launches.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_ -s 8 -c 4 -o $O/r2f_spmm_winner \
    $B > $O/f_ncu_full.log 2>&1
echo "[rc=$?] ncu winner captures" >> $O/f_status.log
unset LGB_SPMM_VARIANT LGB_SPMM_CHUNK
T 400 python tools/shard_probe.py --world 8 --ranks 0 --variants 0,20,23 --chunks 2048,1024,256 > $O/f_shard_probe8.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2f_launches_shard_probe8.csv \
    python tools/shard_probe.py --world 8 --ranks 0 --variants 0,23 --chunks 256,1024 > $O/f_ncu_shard.log 2>&1
T 200 python bench.py --steps 20 --warmup 5 --degree uniform --no-cpu-baseline > $O/f_bench_hm_uniform.json 2> $O/f_bench_hm_uniform.err
cat $O/f_status.log
