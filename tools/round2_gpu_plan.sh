#!/usr/bin/env bash
# GPU measurement plan for the next round.  EVERY command carries its own `timeout`: in round 1 an un-timed 8-rank probe
# hung (NCCL / CUDA-graph capture) and burned 151 GPU-minutes.  Run pieces with:
#   gpurun --timeout 900 -- 'bash tools/round2_gpu_plan.sh single'
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/round2_gpu_plan.sh dist2'
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/round2_gpu_plan.sh dist8'
#   gpurun --timeout 1800 -- 'bash tools/round2_gpu_plan.sh ncu'        (then: python tools/ncu_summary.py gpurun_out/spmm_r2.ncu-rep)
#   gpurun --timeout 1000 -- 'bash tools/round2_gpu_plan.sh sanitize'
set -u
mkdir -p gpurun_out
T() { timeout "$@"; echo "[rc=$?] ${*:2}" | cut -c1-160; }
TR() { local n=$1; shift; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; echo "[rc=$?] torchrun x$n $*" | cut -c1-160; }
case "${1:-single}" in
  single)
    T 600 python -m pytest tests -m gpu -q 2>&1 | tail -15
    T 200 python bench.py --steps 20 --warmup 5 | tail -1 > gpurun_out/bench_hm.json
    T 200 python bench.py --steps 20 --warmup 5 --degree uniform --no-cpu-baseline | tail -1 > gpurun_out/bench_hm_uniform.json
    T 200 python bench.py --steps 50 --warmup 5 --workload ml1m --no-cpu-baseline | tail -1 > gpurun_out/bench_ml1m.json
    T 200 python bench.py --steps 50 --warmup 5 --workload ml1m --no-cpu-baseline --graph | tail -1 > gpurun_out/bench_ml1m_graph.json
    LGB_SPMM_VARIANT=16 T 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline | tail -1 > gpurun_out/bench_hm_v16.json
    T 200 python tools/spmm_probe.py --variants 0,16,18,19,12,13 2>&1 | grep -v Warn
    # one GPU standing in for rank r of an 8- / 4-way sharded run: per-launch latency floor, slice size, CTA-wide slices
    T 400 python tools/shard_probe.py --world 8 --ranks 0,7 --variants 0,16,18,19,12 --chunks 1024,256 2>&1 | grep -v Warn
    T 300 python tools/shard_probe.py --world 4 --ranks 0 --variants 0,16 --chunks 1024,512 2>&1 | grep -v Warn
    for hs in hetero_s hetero_m hetero_l; do
      T 200 python bench.py --workload $hs --steps 20 --warmup 5 | tail -1 > gpurun_out/bench_$hs.json
      T 200 python bench.py --workload $hs --steps 20 --warmup 5 --project-first --no-cpu-baseline | tail -1 > gpurun_out/bench_${hs}_project_first.json
    done
    T 300 python tools/train_lightgcn.py --style reference --iters 100 2>&1 | tail -1 | cut -c1-400
    T 300 python tools/train_lightgcn.py --style fused --iters 100 2>&1 | tail -1 | cut -c1-400
    T 200 python bench.py --steps 20 --warmup 5 --degree-order --no-cpu-baseline | tail -1 > gpurun_out/bench_hm_degree_order.json
    T 600 python tools/sweep.py 2>&1 | tail -12
    T 300 python tools/next_rows_bench.py 2>&1 | grep -v Warn > gpurun_out/next_rows.jsonl
    ;;
  ncu)
    # the top kernel, once: launch list + full capture (B200_PROFILING.md: plain run first, one ncu tool per call)
    # (no T wrapper here: ncu may only run if the plain run of the same command line exited 0)
    B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-autotune"
    timeout 200 $B > gpurun_out/plain_r2.log 2>&1 &&
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv \
        $B > gpurun_out/ncu_launches_r2.log 2>&1 &&
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_ -s 6 -c 4 -o gpurun_out/spmm_r2 \
        $B > gpurun_out/ncu_full_r2.log 2>&1
    echo "[rc=$?] ncu captures"
    ;;
  sanitize)
    # memcheck of the small parity cases (one compute-sanitizer tool per call)
    T 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_lightgcn.py -q -x \
        -k "csr_build or spmm_vs_oracle or lightgcn_against_oracle or bpr_against or topk_against_reference" 2>&1 | tail -15
    ;;
  dist2|dist4|dist8)
    n=${1#dist}
    for mode in "layer,0,nccl" "layer,1,nccl" "pipelined,0,nccl" "pipelined,1,nccl" "merged,0,nccl" "merged,1,nccl" "layer,0,symm" "pipelined,1,symm" "merged,1,symm" "rows,R" "rows,S"; do
      DIST_CHECK_MODE=$mode TR "$n" tests/dist_gpu_check.py 2>&1 | grep -E "DIST_OK|Error|error|rc=" | head -6
    done
    TR "$n" bench.py --gpus "$n" --steps 20 --warmup 5 | tail -2 | cut -c1-300
    LGB_SPMM_VARIANT=16 TR "$n" bench.py --gpus "$n" --steps 20 --warmup 5 | tail -2 | cut -c1-300
    LGB_SPMM_VARIANT=16 TR "$n" bench.py --gpus "$n" --steps 20 --warmup 5 --schedule merged | tail -2 | cut -c1-300
    LGB_SPMM_VARIANT=16 TR "$n" bench.py --gpus "$n" --steps 20 --warmup 5 --schedule merged --exchange symm | tail -2 | cut -c1-300
    TR "$n" bench.py --gpus "$n" --steps 20 --warmup 5 --schedule pipelined | tail -2 | cut -c1-300
    TR "$n" bench.py --gpus "$n" --steps 20 --warmup 5 --schedule pipelined --exchange symm | tail -2 | cut -c1-300
    TR "$n" bench.py --gpus "$n" --steps 20 --warmup 5 --schedule pipelined --graph | tail -2 | cut -c1-300
    PROBE_PIPELINED=1 TR "$n" tools/dist_probe.py 2>&1 | grep -v -i warn | tail -12
    ;;
esac
