#!/bin/bash
# long convergence run on the reference example's observation density (101 observations on the n = 201 grid, phi / sigma fitted)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python tools/cfg5_convergence.py --chains 1024 --obs-every 2 --fit-phi --iters 100000 --leapfrog 100 --diag > gpurun_out/conv_100k.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/conv_100k.log
