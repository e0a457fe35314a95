#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/p_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/p_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/p_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/p_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
print('dense', d['dense']['ms_per_step'], d['dense']['roofline']['frac'])
print('setup', {m:(v['kernel_seconds'], v['roofline']['frac']) for m,v in d['setup']['modes'].items()}, d['setup'].get('evaluation'))
print('cfg5', d['cfg5']['value'], d['cfg5']['sample_seconds'])
PY
