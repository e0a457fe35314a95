#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CHAINS=65536 timeout 600 python tools/hmc_bench.py 2>&1 | tail -1
CHAINS=65536 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 700 --csv --log-file gpurun_out/hmc_launches_r02.csv python tools/hmc_bench.py > gpurun_out/hmc_ncu.log 2>&1; echo "ncu rc=$?"
