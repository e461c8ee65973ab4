set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest_s1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_s1.log
tail -n 5 gpurun_out/r2_pytest_s1.log
timeout 600 python bench.py > gpurun_out/r2_bench_s1.json 2> gpurun_out/r2_bench_s1.err
IGCN_NO_DEFER_DW=1 timeout 600 python bench.py > gpurun_out/r2_bench_s1_nodefer.json 2> gpurun_out/r2_bench_s1_nodefer.err
timeout 600 python bench.py > gpurun_out/r2_bench_s1b.json 2> gpurun_out/r2_bench_s1b.err
timeout 600 python tools/step_timeline.py > gpurun_out/r2_timeline_s1.json 2> gpurun_out/r2_timeline_s1.err
timeout 900 python bench.py --workload config4 > gpurun_out/r2_bench_c4_s1.json 2> gpurun_out/r2_bench_c4_s1.err
python - <<PY
import json
for f in ('bench_s1','bench_s1_nodefer','bench_s1b','bench_c4_s1'):
    try:
        d=json.loads(open('gpurun_out/r2_%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
    except Exception as e: print(f, 'parse', e)
PY
