#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/s_bench$N.json 2> gpurun_out/s_bench$N.err; echo "bench$N rc=$?"; tail -c 800 gpurun_out/s_bench$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/s_bench$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, 'e2e', d['e2e']['value'], d['timing'])
c=d['cfg5']; print({k:c[k] for k in ('value','n_gpus','sample_seconds','allgather_ms','allgather_GBps','draws')})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-900
