#!/bin/bash
# GEMM stage hand-over variants: parity, dense-mode timing per library, setup timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_setup.py tests/test_gpu_baseline_scale.py tests/test_gpu_parity.py tests/test_gpu_initialization.py -x -q 2>&1 | tail -3
for lib in b200 gbk32 gs3d1 gs4d1 gs4d3; do
echo "== $lib"; MAGI_LIB_NAME=libmagi_$lib.so REPS=20 timeout 300 python tools/dense_bench.py 2>&1 | tail -1
done
timeout 600 python bench.py --sections setup,dense > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/v_bench.json').read().strip().splitlines()[-1])
print({k:(v.get('seconds'),v.get('roofline',{}).get('frac')) for k,v in d['setup'].items() if isinstance(v,dict) and 'seconds' in v}, d['dense']['ms_per_step'], d['dense']['roofline']['frac'])
PY
