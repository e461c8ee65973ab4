# quick check after a kernel change: the glue / model / benched parity tests, then the default bench twice
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_glue.py tests/test_gpu_model.py tests/test_gpu_benched.py -m gpu -q -x --timeout 300 > gpurun_out/r2_pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_quick.log
tail -n 3 gpurun_out/r2_pytest_quick.log
timeout 300 python bench.py > gpurun_out/r2_bench_quick.json 2> gpurun_out/r2_bench_quick.err
timeout 300 python tools/step_timeline.py > gpurun_out/r2_timeline_quick.json 2> gpurun_out/r2_timeline_quick.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_quick.json').read().strip().splitlines()[-1])
print('bench', d['value'], d['ms_per_step'], d['e2e']['value'])
k=d['kernels_cupti']
for n,v in k.items():
    if 'lin_bn' in n: print(n, v['us_per_call'], v['calls_per_step'])
PY
