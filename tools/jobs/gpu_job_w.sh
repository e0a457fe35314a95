#!/bin/bash
# dense-mode evaluation: clean run, then the ncu launch list of the same command
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
REPS=20 timeout 300 python tools/dense_bench.py > gpurun_out/w_dense.log 2>&1 || { echo "dense failed"; tail -5 gpurun_out/w_dense.log; exit 1; }
tail -1 gpurun_out/w_dense.log
REPS=3 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/dense_launches_r02.csv python tools/dense_bench.py > gpurun_out/w_ncu.log 2>&1; echo "ncu rc=$?"
