#!/usr/bin/env bash
# Round-2 GPU call E (ONE GPU): parity suite; hot-column kernel after the dynamic work distribution; ranking step with the
# clock sampler started before the warm-up (wgrad kernel on / off); full training iteration with the O(B) device sampler.
set -u
mkdir -p gpurun_out
O=gpurun_out
T() { timeout "$@"; echo "[rc=$?] ${*:2}" | cut -c1-200 >> $O/e_status.log; }
: > $O/e_status.log
T 900 python -m pytest tests -m gpu -q -x --durations=5 > $O/e_pytest.log 2>&1
tail -12 $O/e_pytest.log >> $O/e_status.log
LGB_SPMM_HOT=256 T 200 python tools/spmm_probe.py --variants 0,30 > $O/e_spmm_probe_h256.log 2>&1
LGB_SPMM_HOT=512 T 200 python tools/spmm_probe.py --variants 31 > $O/e_spmm_probe_h512.log 2>&1
LGB_SPMM_HOT=256 T 200 python tools/spmm_probe.py --degree uniform --variants 0,30 > $O/e_spmm_probe_uniform.log 2>&1
LGB_SPMM_HOT=256 T 300 python tools/shard_probe.py --world 8 --ranks 0 --variants 0,30 --chunks 256 > $O/e_shard_probe8_h256.log 2>&1
for hs in hetero_s hetero_m hetero_l; do
  T 200 python bench.py --workload $hs --steps 100 --warmup 10 > $O/e_bench_$hs.json 2> $O/e_bench_$hs.err
  T 200 python bench.py --workload $hs --steps 100 --warmup 10 --hetero-aggr mean --no-cpu-baseline > $O/e_bench_${hs}_mean.json 2> $O/e_bench_${hs}_mean.err
done
LGB_WGRAD_MIN_ROWS=1000000000 T 200 python bench.py --workload hetero_l --steps 100 --warmup 10 --no-cpu-baseline > $O/e_bench_hetero_l_cublas_wgrad.json 2> $O/e_bench_hetero_l_cublas_wgrad.err
LGB_WGRAD_MIN_ROWS=512 T 200 python bench.py --workload hetero_m --steps 100 --warmup 10 --no-cpu-baseline > $O/e_bench_hetero_m_wgrad512.json 2> $O/e_bench_hetero_m_wgrad512.err
T 400 python tools/train_lightgcn.py --workload hm --style fused --sampler device --iters 50 --eval-every 1000 > $O/e_train_hm_device_sampler.log 2>&1
T 300 python tools/train_lightgcn.py --workload ml1m --style fused --sampler device --iters 200 --eval-every 1000 > $O/e_train_ml1m_device_sampler.log 2>&1
T 300 python bench.py --steps 20 --warmup 5 > $O/e_bench_hm.json 2> $O/e_bench_hm.err
T 200 python bench.py --steps 50 --warmup 5 --workload ml1m > $O/e_bench_ml1m.json 2> $O/e_bench_ml1m.err
T 200 python bench.py --steps 50 --warmup 5 --workload ml1m --graph on --no-cpu-baseline > $O/e_bench_ml1m_graph.json 2> $O/e_bench_ml1m_graph.err
cat $O/e_status.log
