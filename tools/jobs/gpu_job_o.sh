#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/o_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/o_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/o_bench2.json 2> gpurun_out/o_bench2.err; echo "bench2 rc=$?"; tail -c 1200 gpurun_out/o_bench2.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/o_bench2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'])
    c=d['cfg5']; print({k:c[k] for k in ('value','n_gpus','sample_seconds','allgather_ms','allgather_GBps','draws','rhat')})
except Exception as e: print('parse failed', e)
PY
