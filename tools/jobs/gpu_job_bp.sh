#!/bin/bash
# one full capture of the Lorenz-96 band-product kernel (cfg4 evaluation) after the evaluation ran clean
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/cfg4_eval.py 2>&1 | tail -1
REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:band_product -s 6 -c 1 -f -o gpurun_out/band_product_r02 python tools/cfg4_eval.py > gpurun_out/bp_ncu.log 2>&1; echo "capture rc=$?"
