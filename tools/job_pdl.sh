set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_pytest_p0.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_p0.log
timeout 400 python bench.py > gpurun_out/r2_bench_p0.json 2> gpurun_out/r2_bench_p0.err
IGCN_PDL=1 timeout 600 python -m pytest tests/test_gpu_benched.py tests/test_gpu_model.py tests/test_gpu_glue.py tests/test_gpu_tc.py -m gpu -q -x --timeout 300 > gpurun_out/r2_pytest_p1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_p1.log
IGCN_PDL=1 timeout 400 python bench.py > gpurun_out/r2_bench_p1.json 2> gpurun_out/r2_bench_p1.err
IGCN_PDL=1 timeout 600 python bench.py --workload config4 > gpurun_out/r2_bench_c4_p1.json 2> gpurun_out/r2_bench_c4_p1.err
tail -n 3 gpurun_out/r2_pytest_p0.log gpurun_out/r2_pytest_p1.log
tail -n 3 gpurun_out/r2_bench_p1.err
python - <<PY
import json
for f in ('p0','p1','c4_p1'):
    try:
        d=json.loads(open('gpurun_out/r2_bench_%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('loss_last'))
    except Exception as e: print(f,'parse',e)
PY
