#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
MAGI_K1=narrow B=2 CHAINS=65536 timeout 600 ncu --set full --clock-control none --import-source on -k regex:narrow_logpost -s 3 -c 1 -f -o gpurun_out/narrow_b2 python tools/quick_bench.py > gpurun_out/n_ncu.log 2>&1; echo "capture rc=$?"
