#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for cfg in "4000 30" "8000 50"; do set -- $cfg
echo "== hmc iters=$1 lf=$2"; timeout 1200 python tools/multi_gpu_hmc.py --chains 1024 --iters $1 --warmup $(($1/2)) --leapfrog $2 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('sample_seconds_max','theta_mean','rhat','ess_bulk_256chains')})"
done
