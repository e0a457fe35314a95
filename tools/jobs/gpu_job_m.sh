#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for cfg in "41 2000 40" "81 2000 40" "41 4000 80"; do set -- $cfg
echo "== n=$1 iters=$2 lf=$3"; timeout 1200 python tools/multi_gpu_hmc.py --n $1 --chains 2048 --iters $2 --warmup $(($2/2)) --leapfrog $3 --step 0.01 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('sample_seconds_max','theta_mean','rhat','ess_bulk_256chains')})"
done
