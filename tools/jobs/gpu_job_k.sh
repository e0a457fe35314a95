#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for it in 300 600; do for lf in 50 100 200; do
echo "== hmc iters=$it lf=$lf"; timeout 900 python tools/multi_gpu_hmc.py --chains 2048 --iters $it --warmup $((it/2)) --leapfrog $lf 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('sample_seconds_max','theta_mean','rhat','ess_bulk_256chains')})"
done; done
