set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_pytest_s.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_s.log
timeout 400 python bench.py > gpurun_out/r2_bench_s.json 2> gpurun_out/r2_bench_s.err
IGCN_NO_AUX_STREAM=1 timeout 400 python bench.py > gpurun_out/r2_bench_s_noaux.json 2> gpurun_out/r2_bench_s_noaux.err
tail -n 3 gpurun_out/r2_pytest_s.log
python - <<PY
import json
for f in ('s','s_noaux'):
    try:
        d=json.loads(open('gpurun_out/r2_bench_%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
    except Exception as e: print(f,'parse', e)
PY
