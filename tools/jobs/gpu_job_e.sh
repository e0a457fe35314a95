#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export MAGI_LIB_NAME=libmagi_fast.so CHAINS=16384
timeout 300 python tools/quick_bench.py > gpurun_out/e_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:flow_logpost -s 5 -c 1 -f -o gpurun_out/flow_v2 python tools/quick_bench.py > gpurun_out/e_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/e_ncu.log; cat gpurun_out/e_plain.log | tail -1
