#!/usr/bin/env bash
# Round-2 GPU call J (ONE GPU): clock sampler through NVML (no nvidia-smi child), ranking step with SAGEConv as one autograd node.
set -u
mkdir -p gpurun_out
O=gpurun_out
T() { timeout "$@"; echo "[rc=$?] ${*:2}" | cut -c1-200 >> $O/j_status.log; }
: > $O/j_status.log
T 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/j_bench_hm.json 2> $O/j_bench_hm.err
for hs in hetero_s hetero_m hetero_l; do
  T 200 python bench.py --workload $hs --steps 100 --warmup 10 --no-cpu-baseline > $O/j_bench_$hs.json 2> $O/j_bench_$hs.err
  LGB_SAGE_FUSED=0 T 200 python bench.py --workload $hs --steps 100 --warmup 10 --no-cpu-baseline > $O/j_bench_${hs}_unfused.json 2> $O/j_bench_${hs}_unfused.err
  T 200 python bench.py --workload $hs --steps 100 --warmup 10 --hetero-aggr mean --no-cpu-baseline > $O/j_bench_${hs}_mean.json 2> $O/j_bench_${hs}_mean.err
done
T 200 python tools/hetero_profile.py --workload hetero_s > $O/j_hetero_profile_s.log 2>&1
T 300 python -m pytest tests/test_gpu_hetero.py -m gpu -q -x > $O/j_pytest_hetero.log 2>&1; tail -2 $O/j_pytest_hetero.log >> $O/j_status.log
cat $O/j_status.log
