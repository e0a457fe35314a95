#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/n_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/n_tests.log
for k1 in default windowed; do
echo "== K1=$k1"; if [ $k1 = default ]; then unset MAGI_K1; else export MAGI_K1=$k1; fi
CHAINS=8,256,1184,2048,2368,4096 timeout 300 python tools/quick_bench.py 2>&1 | tail -6 | cut -c1-100
timeout 120 python tools/single_chain_latency.py 2>&1 | tail -1
done
unset MAGI_K1
timeout 600 python tools/setup_comparators.py > gpurun_out/setup_comparators.json 2> gpurun_out/n_cmp.err; echo "cmp rc=$?"; cat gpurun_out/setup_comparators.json | cut -c1-1500
