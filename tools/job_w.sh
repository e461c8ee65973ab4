set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_glue.py tests/test_gpu_model.py tests/test_gpu_benched.py -m gpu -q -x --timeout 600 > gpurun_out/r2_pytest_w.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_w.log
timeout 400 python bench.py > gpurun_out/r2_bench_w.json 2> gpurun_out/r2_bench_w.err
IGCN_NO_LIN_BN=1 timeout 400 python bench.py > gpurun_out/r2_bench_w_nofuse.json 2> gpurun_out/r2_bench_w_nofuse.err
tail -n 4 gpurun_out/r2_pytest_w.log | cut -c1-250
python - <<PY
import json
for f in ('w','w_nofuse'):
    try:
        d=json.loads(open('gpurun_out/r2_bench_%s.json'%f).read().strip().splitlines()[-1])
        k=d['kernels_cupti']
        print(f, d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), [round(x,3) for x in d['ms_per_step_blocks'][::8]], k.get('lin_bn_act_bwd_pair_kernel',{}).get('us_per_call'), k.get('lin_bn_act_fwd_pair_kernel',{}).get('us_per_call'))
    except Exception as e: print(f,'parse', e)
PY
