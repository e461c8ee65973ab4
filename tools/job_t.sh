set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_pytest_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_t.log
timeout 400 python bench.py --workload config3 > gpurun_out/r2_bench_c3_t.json 2> gpurun_out/r2_bench_c3_t.err
timeout 400 python bench.py > gpurun_out/r2_bench_t.json 2> gpurun_out/r2_bench_t.err
tail -n 3 gpurun_out/r2_pytest_t.log
python - <<PY
import json
for f in ('t','c3_t'):
    try:
        d=json.loads(open('gpurun_out/r2_bench_%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d['config'].get('ms_per_step_eager'))
    except Exception as e: print(f,'parse', e)
PY
