#!/usr/bin/env bash
# Round-2 scaling points: bench.py at N GPUs (+ the ranking step as data-parallel replicas at N = 2).
set -u
N=${1:-2}
O=gpurun_out
mkdir -p $O
S=$O/s${N}_status.log
: > $S
TR() { local t=$1; shift; timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; local rc=$?; echo "[rc=$rc] torchrun x$N $*" | cut -c1-200 >> $S; return $rc; }
TR 300 bench.py --gpus "$N" --steps 20 --warmup 5 > $O/s${N}_bench.json 2> $O/s${N}_bench.err
tail -c 600 $O/s${N}_bench.json >> $S
if [ "$N" = "2" ]; then
  TR 200 bench.py --gpus 2 --workload hetero_m --steps 50 --warmup 10 --no-cpu-baseline > $O/s2_bench_hetero_m.json 2> $O/s2_bench_hetero_m.err
  TR 200 bench.py --gpus 2 --workload hetero_l --steps 50 --warmup 10 --no-cpu-baseline > $O/s2_bench_hetero_l.json 2> $O/s2_bench_hetero_l.err
fi
cat $S
