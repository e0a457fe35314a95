#!/bin/bash
# round-2 GPU job A: K1 variants parity + quick bench of both kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_k1_variants.py -x -q > gpurun_out/a_k1tests.log 2>&1; echo "k1tests rc=$?" 
tail -5 gpurun_out/a_k1tests.log
MAGI_K1=flow timeout 300 python tools/quick_bench.py > gpurun_out/a_quick_flow.log 2>&1; echo "flow rc=$?"; cat gpurun_out/a_quick_flow.log | tail -4
MAGI_K1=windowed timeout 300 python tools/quick_bench.py > gpurun_out/a_quick_win.log 2>&1; echo "win rc=$?"; cat gpurun_out/a_quick_win.log | tail -4
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/a_alltests.log 2>&1; echo "alltests rc=$?"
tail -5 gpurun_out/a_alltests.log
