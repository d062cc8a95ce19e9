#!/usr/bin/env bash
# Round-2 GPU call C (ONE GPU): parity suite incl. the full-size fp64 check, headline bench after the long-row reduce change,
# the hetero step after the host-sync removal + a torch.profiler breakdown of it, the 8-way shard probe with both SpMM
# launches of a layer on two streams.
set -u
mkdir -p gpurun_out
O=gpurun_out
T() { timeout "$@"; echo "[rc=$?] ${*:2}" | cut -c1-200 >> $O/c_status.log; }
: > $O/c_status.log
T 900 python -m pytest tests -m gpu -q -x --durations=5 > $O/c_pytest.log 2>&1
tail -12 $O/c_pytest.log >> $O/c_status.log
T 400 python bench.py --steps 20 --warmup 5 > $O/c_bench_hm.json 2> $O/c_bench_hm.err
for hs in hetero_s hetero_m hetero_l; do
  T 200 python bench.py --workload $hs --steps 20 --warmup 5 > $O/c_bench_$hs.json 2> $O/c_bench_$hs.err
  T 200 python bench.py --workload $hs --steps 20 --warmup 5 --hetero-aggr mean --no-cpu-baseline > $O/c_bench_${hs}_mean.json 2> $O/c_bench_${hs}_mean.err
done
T 200 python tools/hetero_profile.py --workload hetero_s > $O/c_hetero_profile_s.log 2>&1
T 200 python tools/hetero_profile.py --workload hetero_l > $O/c_hetero_profile_l.log 2>&1
T 400 python tools/shard_probe.py --world 8 --ranks 0 --variants 0,20,23 --chunks 1024,256 > $O/c_shard_probe8.log 2>&1
T 300 python tools/shard_probe.py --world 4 --ranks 0 --variants 0,20 --chunks 1024,256 > $O/c_shard_probe4.log 2>&1
T 200 python tools/spmm_probe.py --variants 0,20,23 > $O/c_spmm_probe.log 2>&1
cat $O/c_status.log
