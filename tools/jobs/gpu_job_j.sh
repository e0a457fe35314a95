#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/j_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/j_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/j_bench.err; head -c 6000 gpurun_out/j_bench.json
for it in 200 400 800; do for lf in 10 20; do
echo "== hmc iters=$it lf=$lf"; timeout 600 python tools/multi_gpu_hmc.py --chains 4096 --iters $it --warmup $((it/2)) --leapfrog $lf 2>&1 | tail -1 | cut -c1-600
done; done
