#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export MAGI_LIB_NAME=libmagi_fast.so CHAINS=4096
for lag in -1 0 2 4; do
  echo "== order lag=$lag"
  MAGI_FLOW_LAG=$lag MAGI_DBG_CLOCKS=1 timeout 300 python tools/quick_bench.py 2>&1 | grep -E "dbg flow|\"n\"" | tail -2
done
timeout 300 python tools/quick_bench.py > gpurun_out/b_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:flow_logpost -s 5 -c 1 -f -o gpurun_out/flow_v1 python tools/quick_bench.py > gpurun_out/b_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/b_ncu.log
