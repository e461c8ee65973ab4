set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_pytest_u.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_u.log
timeout 400 python bench.py > gpurun_out/r2_bench_u.json 2> gpurun_out/r2_bench_u.err
IGCN_GO_BRANCH=0 timeout 400 python bench.py > gpurun_out/r2_bench_u_nobranch.json 2> gpurun_out/r2_bench_u_nobranch.err
timeout 600 python bench.py --workload config4 > gpurun_out/r2_bench_c4_u.json 2> gpurun_out/r2_bench_c4_u.err
tail -n 3 gpurun_out/r2_pytest_u.log
python - <<PY
import json
for f in ('u','u_nobranch','c4_u'):
    try:
        d=json.loads(open('gpurun_out/r2_bench_%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), [round(x,3) for x in d['ms_per_step_blocks'][::6]])
    except Exception as e: print(f,'parse', e)
PY
