#!/usr/bin/env bash
# Round-2 GPU call A (ONE GPU): the measurements round 1 never made.  Every command carries its own timeout.
#   gpurun --timeout 1500 -- 'bash tools/r2_call_a.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
T() { timeout "$@"; echo "[rc=$?] ${*:2}" | cut -c1-200 >> $O/a_status.log; }
: > $O/a_status.log
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/a_gpu.txt 2>&1

# 1. headline bench (autotune on) -> winner configuration
T 400 python bench.py --steps 20 --warmup 5 > $O/a_bench_hm.json 2> $O/a_bench_hm.err
eval "$(python tools/pick_winner.py $O/a_bench_hm.json)"     # FWD_V FWD_CHUNK FWD_ORDER (forward graph's plan)
echo "winner: v=$FWD_V chunk=$FWD_CHUNK order=$FWD_ORDER" >> $O/a_status.log
ORD=""; [ "$FWD_ORDER" = "1" ] && ORD="--degree-order"

# 2. ncu of the kernel that ships: plain run first, then launch list, then one full capture (B200_PROFILING.md)
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-autotune $ORD"
export LGB_SPMM_VARIANT=$FWD_V LGB_SPMM_CHUNK=$FWD_CHUNK
timeout 200 $B > $O/a_plain_winner.json 2> $O/a_plain_winner.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2a_launches_bench_hm.csv \
    $B > $O/a_ncu_launches.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_ -s 8 -c 4 -o $O/r2a_spmm_winner \
    $B > $O/a_ncu_full.log 2>&1
echo "[rc=$?] ncu winner captures" >> $O/a_status.log
# the default variant (0, chunk 1024) for comparison: one full capture
export LGB_SPMM_VARIANT=0 LGB_SPMM_CHUNK=1024
B0="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-autotune"
timeout 200 $B0 > $O/a_plain_v0.json 2> $O/a_plain_v0.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_subwarp -s 8 -c 2 -o $O/r2a_spmm_v0 \
    $B0 > $O/a_ncu_full_v0.log 2>&1
echo "[rc=$?] ncu v0 capture" >> $O/a_status.log
unset LGB_SPMM_VARIANT LGB_SPMM_CHUNK

# 3. the other BASELINE.json configs
T 200 python bench.py --steps 20 --warmup 5 --degree uniform --no-cpu-baseline > $O/a_bench_hm_uniform.json 2> $O/a_bench_hm_uniform.err
T 200 python bench.py --steps 50 --warmup 5 --workload ml1m > $O/a_bench_ml1m.json 2> $O/a_bench_ml1m.err
T 200 python bench.py --steps 50 --warmup 5 --workload ml1m --no-cpu-baseline --graph > $O/a_bench_ml1m_graph.json 2> $O/a_bench_ml1m_graph.err
for hs in hetero_s hetero_m hetero_l; do
  T 200 python bench.py --workload $hs --steps 20 --warmup 5 > $O/a_bench_$hs.json 2> $O/a_bench_$hs.err
  T 200 python bench.py --workload $hs --steps 20 --warmup 5 --hetero-aggr mean --no-cpu-baseline > $O/a_bench_${hs}_mean.json 2> $O/a_bench_${hs}_mean.err
done
T 300 python tools/train_lightgcn.py --style reference --iters 30 > $O/a_train_reference.log 2>&1
T 300 python tools/train_lightgcn.py --style fused --iters 30 > $O/a_train_fused.log 2>&1
T 500 python tools/train_lightgcn.py --workload hm --style both --iters 10 --eval-every 1000 > $O/a_train_hm.log 2>&1
T 300 python tools/next_rows_bench.py > $O/a_next_rows.jsonl 2> $O/a_next_rows.err
T 700 python tools/sweep.py --steps 10 > $O/a_sweep_1gpu.log 2>&1

# 4. one GPU standing in for rank r of an 8-way sharded run: per-launch floor
T 400 python tools/shard_probe.py --world 8 --ranks 0,5 --variants 0,16,20,23,25 --chunks 1024,256 > $O/a_shard_probe8.log 2>&1
T 200 python tools/spmm_probe.py --variants 0,20,23,25 > $O/a_spmm_probe.log 2>&1

# 5. GPU parity tests last (unchanged code since the driver's round-1 run)
T 600 python -m pytest tests -m gpu -q -x > $O/a_pytest.log 2>&1
tail -3 $O/a_pytest.log >> $O/a_status.log
cat $O/a_status.log
