#!/usr/bin/env bash
# Round-2 final single-GPU validation: parity suite, smoke, the bench lines quoted in BASELINE.md.
set -u
mkdir -p gpurun_out
O=gpurun_out
T() { timeout "$@"; echo "[rc=$?] ${*:2}" | cut -c1-200 >> $O/i_status.log; }
: > $O/i_status.log
T 900 python -m pytest tests -m gpu -q -x --durations=5 > $O/i_pytest.log 2>&1
tail -8 $O/i_pytest.log >> $O/i_status.log
T 300 python __graft_entry__.py --smoke > $O/i_smoke.log 2>&1; tail -1 $O/i_smoke.log >> $O/i_status.log
T 500 python bench.py --steps 20 --warmup 5 > $O/i_bench_hm.json 2> $O/i_bench_hm.err
T 500 python bench.py --impl reference --steps 5 --warmup 1 > $O/i_bench_hm_reference.json 2> $O/i_bench_hm_reference.err
T 200 python bench.py --steps 50 --warmup 5 --workload ml1m > $O/i_bench_ml1m.json 2> $O/i_bench_ml1m.err
for hs in hetero_s hetero_m hetero_l; do
  T 200 python bench.py --workload $hs --steps 100 --warmup 10 > $O/i_bench_$hs.json 2> $O/i_bench_$hs.err
done
T 200 python bench.py --impl reference --workload hetero_s --steps 5 --warmup 1 > $O/i_bench_hetero_reference.json 2> $O/i_bench_hetero_reference.err
T 300 python tools/train_lightgcn.py --workload hm --style both --iters 10 --eval-every 1000 > $O/i_train_hm_reference_sampler.log 2>&1
T 300 python tools/train_lightgcn.py --workload hm --style fused --sampler device --iters 100 --eval-every 1000 > $O/i_train_hm_device_sampler.log 2>&1
T 300 python tools/next_rows_bench.py > $O/i_next_rows.jsonl 2> $O/i_next_rows.err
cat $O/i_status.log
