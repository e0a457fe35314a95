#!/bin/bash
# K1-narrow: parity, then timing at band half-widths 4, 2, 1
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_k1_variants.py -x -q -m gpu -k "narrow" > gpurun_out/n_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/n_tests.log
for b in 4 2 1; do
echo "== b=$b narrow"; MAGI_K1=narrow B=$b CHAINS=4096,8192,16384,32768,65536 timeout 300 python tools/quick_bench.py 2>&1 | tail -5 | cut -c1-100
done
