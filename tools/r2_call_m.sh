#!/usr/bin/env bash
# Round-2 GPU call M (ONE GPU): column-sweep order at smaller slice sizes; the decoder's node-projection inference form.
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/m_status.log
run() { local name=$1; shift; timeout "$1" "${@:2}" > $O/$name.log 2>&1; echo "[rc=$?] ${*:2}" >> $O/m_status.log; }
run m_pytest 300 python -m pytest tests/test_gpu_hetero.py -q -x -k "decoder or infer"
run m_probe_sweep 500 python tools/spmm_probe.py --variants 0,20,23,24 --chunks 512,256,128,64 --sweep 1
run m_probe_nosweep 300 python tools/spmm_probe.py --variants 0,23 --chunks 256,128 --sweep 0
run m_probe_uniform 300 python tools/spmm_probe.py --degree uniform --variants 0,23 --chunks 1024,256 --sweep 0,1
cat $O/m_status.log; grep -h "^chunk" $O/m_probe_sweep.log $O/m_probe_nosweep.log; echo uniform; grep -h "^chunk\|^graph" $O/m_probe_uniform.log; tail -3 $O/m_pytest.log
