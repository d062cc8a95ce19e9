#!/usr/bin/env bash
# Round-2 multi-GPU call: hardware parity of the sharded engine in every configuration the bench may select, component
# probe, bench.  EVERY torchrun carries its own timeout (a barrier that cannot complete traps after ~4 s by itself).
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_call_b.sh 2'
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/r2_call_b.sh 8 quick'
set -u
N=${1:-2}
QUICK=${2:-}
O=gpurun_out
mkdir -p $O
S=$O/b${N}_status.log
: > $S
TR() { local t=$1; shift; timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; local rc=$?; echo "[rc=$rc] torchrun x$N $*" | cut -c1-200 >> $S; return $rc; }
nvidia-smi topo -m > $O/b${N}_topo.txt 2>&1

if [ -z "$QUICK" ]; then
  MODES="chains,auto chains,symm,graph layer,symm chains,nccl layer,nccl rows,R rows,S"
else
  MODES="chains,auto chains,symm,graph"
fi
for mode in $MODES; do
  DIST_CHECK_MODE=$mode TR 150 tests/dist_gpu_check.py > $O/b${N}_check_${mode//,/_}.log 2>&1
  grep -h "DIST_OK" $O/b${N}_check_${mode//,/_}.log | sort >> $S
  grep -h -i "error\|timed out\|Traceback" $O/b${N}_check_${mode//,/_}.log | head -5 >> $S
done
cat $O/b${N}_check_*.log | grep DIST_OK | sort > $O/r2_dist_ok_${N}gpu.log

NCCL_DEBUG=WARN TR 240 tools/dist_probe.py > $O/b${N}_probe.log 2>&1
grep -v -i warn $O/b${N}_probe.log | tail -16 >> $S
TR 300 bench.py --gpus "$N" --steps 20 --warmup 5 > $O/b${N}_bench.json 2> $O/b${N}_bench.err
tail -c 1500 $O/b${N}_bench.json >> $S
if [ -z "$QUICK" ]; then
  TR 300 bench.py --gpus "$N" --steps 20 --warmup 5 --exchange nccl --graph off > $O/b${N}_bench_nccl.json 2> $O/b${N}_bench_nccl.err
  TR 300 bench.py --gpus "$N" --steps 20 --warmup 5 --graph off --no-autotune > $O/b${N}_bench_nograph_noautotune.json 2> $O/b${N}_bench_nograph_noautotune.err
  timeout 300 python -m pytest tests -m gpu -q -x > $O/b${N}_pytest.log 2>&1; tail -3 $O/b${N}_pytest.log >> $S
fi
cat $S
