#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export MAGI_LIB_NAME=libmagi_fast.so
for st in 0 300 700 1200 2000; do for lag in -1 3; do
echo "== stagger=$st lag=$lag"; MAGI_FLOW_STAGGER=$st MAGI_FLOW_LAG=$lag CHAINS=4096,65536 timeout 300 python tools/quick_bench.py 2>&1 | tail -2 | cut -c1-100
done; done
