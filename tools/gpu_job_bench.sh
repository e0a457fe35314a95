#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SECONDS=0
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$? in ${SECONDS}s"; tail -c 400 gpurun_out/f_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/f_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['timing']['ms_per_step_min'], d['timing']['ms_per_step_max'])
print('dense', d['dense']['ms_per_step'], d['dense']['roofline']['frac'], d['dense']['gpu_launches'])
print('setup', {m: (v['kernel_seconds'], v['roofline']['frac']) for m, v in d['setup']['modes'].items()}, 'eval', d['setup']['evaluation']['ms_per_step'], d['setup']['evaluation']['roofline']['frac'])
print('cfg5', d['cfg5']['value'], d['cfg5'].get('sample_seconds'))
PY
