#!/usr/bin/env bash
# Round-2 GPU call L (ONE GPU): column-sweep order of the long-row slices (lgb_csr.task_exec) -- parity, half-launch probe,
# bench with the autotune choosing; and the ranking step with gc.freeze / default / disable.
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/l_status.log
run() { local name=$1; shift; timeout "$1" "${@:2}" > $O/$name.log 2>&1; echo "[rc=$?] ${*:2}" >> $O/l_status.log; }
run l_pytest_sweep 300 python -m pytest tests/test_gpu_lightgcn.py -q -x -k "sweep or degree_order or large"
run l_probe_sweep 400 python tools/spmm_probe.py --variants 0,23 --chunks 4096,1024,512 --sweep 0,1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/l_bench_hm.json 2> $O/l_bench_hm.err; echo "[rc=$?] bench hm" >> $O/l_status.log
for mode in freeze default disable; do
  for hs in hetero_s hetero_l; do
    LGB_BENCH_GC=$mode timeout 200 python bench.py --workload $hs --steps 200 --warmup 10 --no-cpu-baseline > $O/l_bench_${hs}_$mode.json 2> $O/l_bench_${hs}_$mode.err
    python - $O/l_bench_${hs}_$mode.json >> $O/l_status.log <<'PY'
import json,sys
try:
    j=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    print(sys.argv[1], 'mean', round(j['ms_per_step'],3), {k:(round(v,3) if isinstance(v,float) else v) for k,v in j['per_step_ms'].items() if k!='note'}, 'e2e', round(j['e2e']['ms_per_step'],3))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
  done
done
python - $O/l_bench_hm.json >> $O/l_status.log <<'PY'
import json,sys
j=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print('hm ms', j['ms_per_step'], 'e2e', j['e2e'], 'chosen', j['config']['spmm_variant'].get('forward'), j['config']['spmm_variant'].get('backward'))
print('roofline', {k:v for k,v in j['roofline'].items() if k in ('achieved','frac','traffic','dram_frac')})
PY
cat $O/l_status.log; tail -30 $O/l_probe_sweep.log
