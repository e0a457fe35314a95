#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for lib in fast w20 w24; do for lag in -1 3; do
echo "== lib=$lib lag=$lag"; MAGI_LIB_NAME=libmagi_$lib.so MAGI_FLOW_LAG=$lag CHAINS=4096,65536 timeout 300 python tools/quick_bench.py 2>&1 | tail -2 | cut -c1-100
done; done
MAGI_LIB_NAME=libmagi_w20.so timeout 600 python -m pytest tests/test_gpu_k1_variants.py -x -q 2>&1 | tail -2
