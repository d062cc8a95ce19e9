#!/usr/bin/env bash
# Round-2 GPU call K (ONE GPU): where does the run-to-run variance of the (host-bound) ranking step come from?
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/k_status.log
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" > $O/k_lscpu.txt 2>&1
cat /sys/bus/pci/devices/*/local_cpulist 2>/dev/null | sort | uniq -c | sort -rn | head -5 >> $O/k_lscpu.txt
python - >> $O/k_lscpu.txt 2>&1 <<'PY'
import os; print("affinity", sorted(os.sched_getaffinity(0)))
PY
for rep in 1 2 3 4; do
  for hs in hetero_s hetero_l; do
    timeout 200 python bench.py --workload $hs --steps 100 --warmup 10 --no-cpu-baseline > $O/k_bench_${hs}_$rep.json 2> $O/k_bench_${hs}_$rep.err
    python - $O/k_bench_${hs}_$rep.json >> $O/k_status.log <<'PY'
import json,sys
j=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print(sys.argv[1], 'mean', round(j['ms_per_step'],3), j['per_step_ms'] and {k:round(v,3) for k,v in j['per_step_ms'].items() if k!='note'}, 'e2e', round(j['e2e']['ms_per_step'],3), j['config'].get('cpu_affinity'))
PY
  done
done
timeout 200 python bench.py --steps 50 --warmup 5 --workload ml1m --no-cpu-baseline > $O/k_bench_ml1m.json 2> $O/k_bench_ml1m.err
cat $O/k_status.log; cat $O/k_lscpu.txt
