#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export MAGI_LIB_NAME=libmagi_fast.so
timeout 900 python -m pytest tests/test_gpu_k1_variants.py -x -q 2>&1 | tail -3
CHAINS=4096,65536 timeout 300 python tools/quick_bench.py 2>&1 | tail -2
CHAINS=4096 MAGI_DBG_CLOCKS=1 timeout 300 python tools/quick_bench.py 2>&1 | grep "dbg flow" | tail -1
MAGI_LIB_NAME=libmagi_tl.so CHAINS=4096 MAGI_DBG_CLOCKS=1 timeout 300 python tools/quick_bench.py 2>&1 | grep "dbg flow" | tail -2
MAGI_LIB_NAME=libmagi_tl.so CHAINS=65536 MAGI_DBG_CLOCKS=1 timeout 300 python tools/quick_bench.py 2>&1 | grep "dbg flow" | tail -2
