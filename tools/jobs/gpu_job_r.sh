#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --sections dense,setup > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
print('dense', d['dense']['ms_per_step'], d['dense']['roofline']['frac'], d['dense']['gpu_launches'])
e=d['setup']['evaluation']; print('cfg4 eval', e['ms_per_step'], e['roofline']['frac'], e['gpu_launches'])
PY
python tools/cfg4_eval.py > gpurun_out/r_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"band_product|dense_" --csv --log-file gpurun_out/r_launches.csv python tools/cfg4_eval.py > gpurun_out/r_ncu.log 2>&1; tail -8 gpurun_out/r_launches.csv | awk -F'","' '{print $5, $NF}'
