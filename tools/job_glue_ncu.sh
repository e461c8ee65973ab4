set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_glue.py tests/test_gpu_model.py tests/test_gpu_benched.py -m gpu -q -x --timeout 500 > gpurun_out/r2_pytest_r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_r.log
timeout 400 python bench.py > gpurun_out/r2_bench_r.json 2> gpurun_out/r2_bench_r.err
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:bn_act|skinny|go_spmm' --launch-skip 40 -c 20 -f -o gpurun_out/r2_glue_c2 python tools/prof_kernels.py --compact --iters 2 --what go --B 512 > gpurun_out/r2_glue_ncu.log 2>&1
tail -n 3 gpurun_out/r2_pytest_r.log
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_r.json').read().strip().splitlines()[-1])
    print('bench', d['value'], d['ms_per_step'], d['e2e']['value'])
except Exception as e: print('parse', e)
PY
ls -la gpurun_out/
