#!/bin/bash
# end-of-round validation: smoke, GPU tests, both bench arms, then (each only after its command ran clean) the ncu launch list of the
# bench command and one full capture of the two hot kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/f_tests.log
SECONDS=0
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$? in ${SECONDS}s"; tail -c 400 gpurun_out/f_bench.err
SECONDS=0
timeout 900 python bench.py --impl reference > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "reference arm rc=$? in ${SECONDS}s"; tail -c 300 gpurun_out/f_bench_ref.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/f_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['timing']['ms_per_step_min'], d['timing']['ms_per_step_max'])
print('dense', d['dense']['ms_per_step'], d['dense']['roofline']['frac'], d['dense']['gpu_launches'])
print('setup', {m: (v['kernel_seconds'], v['roofline']['frac']) for m, v in d['setup']['modes'].items()}, 'eval', d['setup']['evaluation']['ms_per_step'], d['setup']['evaluation']['roofline']['frac'])
print('cfg5', d['cfg5']['value'], d['cfg5'].get('sample_seconds'))
PY
timeout 600 python bench.py --steps 5 --warmup 3 --repeats 2 --sections none --no-cpu-baseline > gpurun_out/f_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 5 --warmup 3 --repeats 2 --sections none --no-cpu-baseline > gpurun_out/f_ncu1.log 2>&1; echo "launch list rc=$?"
CHAINS=4096 MAGI_K1=windowed timeout 600 ncu --set full --clock-control none --import-source on -k regex:banded_logpost -s 5 -c 1 -f -o gpurun_out/k1_r02 python tools/quick_bench.py > gpurun_out/f_ncu2.log 2>&1; echo "k1 capture rc=$?"
REPS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_dmma_streamk -s 8 -c 1 -f -o gpurun_out/k2_r02 python tools/dense_bench.py > gpurun_out/f_ncu3.log 2>&1; echo "k2 capture rc=$?"
ls -la gpurun_out/*.ncu-rep
