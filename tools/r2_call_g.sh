#!/usr/bin/env bash
# Round-2 GPU call G (ONE GPU): stage 2 as a tree (one warp per 32 partial rows, tickets) vs flat, at N = 1 and on a 1/8 shard.
set -u
mkdir -p gpurun_out
O=gpurun_out
T() { timeout "$@"; echo "[rc=$?] ${*:2}" | cut -c1-200 >> $O/g_status.log; }
: > $O/g_status.log
T 900 python -m pytest tests -m gpu -q -x --durations=5 > $O/g_pytest.log 2>&1
tail -8 $O/g_pytest.log >> $O/g_status.log
T 500 python bench.py --steps 20 --warmup 5 > $O/g_bench_hm.json 2> $O/g_bench_hm.err
LGB_SPMM_STAGE2=flat T 500 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/g_bench_hm_flat_stage2.json 2> $O/g_bench_hm_flat_stage2.err
eval "$(python tools/pick_winner.py $O/g_bench_hm.json)"
echo "winner: v=$FWD_V chunk=$FWD_CHUNK order=$FWD_ORDER hot=$FWD_HOT" >> $O/g_status.log
ORD=""; [ "$FWD_ORDER" = "1" ] && ORD="--degree-order"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-autotune $ORD"
export LGB_SPMM_VARIANT=$FWD_V LGB_SPMM_CHUNK=$FWD_CHUNK
timeout 200 $B > $O/g_plain_winner.json 2> $O/g_plain_winner.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2g_launches_bench_hm.csv \
    $B > $O/g_ncu_launches.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_ -s 8 -c 4 -o $O/r2g_spmm_winner \
    $B > $O/g_ncu_full.log 2>&1
echo "[rc=$?] ncu winner captures" >> $O/g_status.log
unset LGB_SPMM_VARIANT LGB_SPMM_CHUNK
T 400 python tools/shard_probe.py --world 8 --ranks 0 --variants 0,20,23 --chunks 1024,512,256 > $O/g_shard_probe8.log 2>&1
LGB_SPMM_STAGE2=flat T 400 python tools/shard_probe.py --world 8 --ranks 0 --variants 0 --chunks 256 > $O/g_shard_probe8_flat.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2g_launches_shard_probe8.csv \
    python tools/shard_probe.py --world 8 --ranks 0 --variants 0 --chunks 256 > $O/g_ncu_shard.log 2>&1
T 300 python tools/train_lightgcn.py --workload hm --style fused --sampler device --iters 50 --eval-every 1000 > $O/g_train_hm_device_sampler.log 2>&1
cat $O/g_status.log
