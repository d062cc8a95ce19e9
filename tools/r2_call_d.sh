#!/usr/bin/env bash
# Round-2 GPU call D (ONE GPU): the hot-column-cache SpMM (variants 30 / 31) -- parity suite, half-by-half probe, headline
# bench with the plan-time choice, ncu of the winner, the 8-way shard probe.
set -u
mkdir -p gpurun_out
O=gpurun_out
T() { timeout "$@"; echo "[rc=$?] ${*:2}" | cut -c1-200 >> $O/d_status.log; }
: > $O/d_status.log
T 900 python -m pytest tests -m gpu -q -x --durations=5 > $O/d_pytest.log 2>&1
tail -12 $O/d_pytest.log >> $O/d_status.log
LGB_SPMM_HOT=256 T 200 python tools/spmm_probe.py --variants 0,20,30 > $O/d_spmm_probe_h256.log 2>&1
LGB_SPMM_HOT=512 T 200 python tools/spmm_probe.py --variants 31 > $O/d_spmm_probe_h512.log 2>&1
LGB_SPMM_HOT=768 T 200 python tools/spmm_probe.py --variants 31 > $O/d_spmm_probe_h768.log 2>&1
LGB_SPMM_HOT=256 T 200 python tools/spmm_probe.py --degree uniform --variants 0,30 > $O/d_spmm_probe_uniform.log 2>&1
T 500 python bench.py --steps 20 --warmup 5 > $O/d_bench_hm.json 2> $O/d_bench_hm.err
eval "$(python tools/pick_winner.py $O/d_bench_hm.json)"
echo "winner: v=$FWD_V chunk=$FWD_CHUNK order=$FWD_ORDER hot=$FWD_HOT" >> $O/d_status.log
ORD=""; [ "$FWD_ORDER" = "1" ] && ORD="--degree-order"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-autotune $ORD"
export LGB_SPMM_VARIANT=$FWD_V LGB_SPMM_CHUNK=$FWD_CHUNK LGB_SPMM_HOT=$FWD_HOT
timeout 200 $B > $O/d_plain_winner.json 2> $O/d_plain_winner.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2d_launches_bench_hm.csv \
    $B > $O/d_ncu_launches.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_ -s 8 -c 4 -o $O/r2d_spmm_winner \
    $B > $O/d_ncu_full.log 2>&1
echo "[rc=$?] ncu winner captures" >> $O/d_status.log
unset LGB_SPMM_VARIANT LGB_SPMM_CHUNK LGB_SPMM_HOT
LGB_SPMM_HOT=256 T 400 python tools/shard_probe.py --world 8 --ranks 0 --variants 0,30 --chunks 1024,256 > $O/d_shard_probe8_h256.log 2>&1
LGB_SPMM_HOT=512 T 400 python tools/shard_probe.py --world 8 --ranks 0 --variants 31 --chunks 256 > $O/d_shard_probe8_h512.log 2>&1
for hs in hetero_s hetero_m hetero_l; do
  T 200 python bench.py --workload $hs --steps 20 --warmup 5 --no-cpu-baseline > $O/d_bench_$hs.json 2> $O/d_bench_$hs.err
  LGB_WGRAD_MIN_ROWS=1000000000 T 200 python bench.py --workload $hs --steps 20 --warmup 5 --no-cpu-baseline > $O/d_bench_${hs}_cublas_wgrad.json 2> $O/d_bench_${hs}_cublas_wgrad.err
done
T 200 python tools/hetero_profile.py --workload hetero_l > $O/d_hetero_profile_l.log 2>&1
cat $O/d_status.log
