set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_pytest_v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_v.log
timeout 400 python bench.py > gpurun_out/r2_bench_v.json 2> gpurun_out/r2_bench_v.err
IGCN_NO_LIN_BN=1 timeout 400 python bench.py > gpurun_out/r2_bench_v_nofuse.json 2> gpurun_out/r2_bench_v_nofuse.err
tail -n 30 gpurun_out/r2_pytest_v.log | cut -c1-250
python - <<PY
import json
for f in ('v','v_nofuse'):
    try:
        d=json.loads(open('gpurun_out/r2_bench_%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), [round(x,3) for x in d['ms_per_step_blocks'][::8]])
    except Exception as e: print(f,'parse', e)
PY
