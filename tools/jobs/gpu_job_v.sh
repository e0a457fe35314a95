#!/bin/bash
# dense-mode / setup check after GEMM changes: parity, dense timing, setup + dense bench sections
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_setup.py tests/test_gpu_baseline_scale.py tests/test_gpu_parity.py tests/test_gpu_initialization.py -x -q 2>&1 | tail -3
REPS=20 timeout 300 python tools/dense_bench.py 2>&1 | tail -1
timeout 600 python bench.py --sections setup,dense > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/v_bench.json').read().strip().splitlines()[-1])
print({k:(v['kernel_seconds'],v['roofline']['frac']) for k,v in d['setup']['modes'].items()}, d['dense']['ms_per_step'], d['dense']['roofline']['frac'], d['dense']['gpu_launches'])
PY
