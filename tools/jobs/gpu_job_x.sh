#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() { echo "== $*"; timeout 900 python tools/cfg5_convergence.py "$@" 2>&1 | tail -2; }
run --chains 1024 --n 397 --obs-every 4 --beta 1,1,5 --fit-phi --iters 600 --leapfrog 20 --diag
run --chains 1024 --n 397 --obs-every 4 --beta 1,1,1 --fit-phi --iters 600 --leapfrog 20 --diag
run --chains 1 --n 397 --obs-every 4 --beta 1,1,5 --fit-phi --iters 600 --leapfrog 20 --diag
run --chains 1024 --n 397 --obs-every 4 --beta 1,1,5 --fit-phi --iters 600 --depth 6 --diag
