#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/u_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/u_tests.log
timeout 900 python bench.py > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/u_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/u_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['timing']['ms_per_step_min'], d['timing']['ms_per_step_max'])
print('dense', d['dense']['ms_per_step'], d['dense']['roofline']['frac'], d['dense']['gpu_launches'])
print('cfg5', d['cfg5']['value'], 'aligned', d['cfg5_wave_aligned']['value'], d['cfg5_wave_aligned']['config']['chains_total'])
PY
