set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_tc.py -m gpu -q -x --timeout 500 > gpurun_out/r2_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > gpurun_out/r2_bench_2gpu_v2.json 2> gpurun_out/r2_bench_2gpu_v2.err
tail -n 4 gpurun_out/r2_pytest_2gpu.log
tail -n 3 gpurun_out/r2_bench_2gpu_v2.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_2gpu_v2.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('replicas_identical'), d.get('dp_check'), d.get('config4_dp'))
except Exception as e: print('parse', e)
PY
