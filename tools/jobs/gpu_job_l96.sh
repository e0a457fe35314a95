#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_scale.py -x -q -m gpu -k "lorenz96 or config4" 2>&1 | tail -2
timeout 600 python tools/cfg4_eval.py 2>&1 | tail -1
REPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"band_product|dense_|lorenz96" --csv --log-file gpurun_out/cfg4_eval_launches_r02b.csv python tools/cfg4_eval.py > gpurun_out/l96_ncu.log 2>&1; echo "ncu rc=$?"
