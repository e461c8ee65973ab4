# config-2 bench line at 8 GPUs (the scaling run's largest point): fused all-reduce + Adam checked against ncclAllReduce inside bench.py
set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 > gpurun_out/r2_bench_config2_8gpu.json 2> gpurun_out/r2_bench_config2_8gpu.err
tail -n 3 gpurun_out/r2_bench_config2_8gpu.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_config2_8gpu.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('dp_check'), d['config'].get('replicas_identical'), d.get('config4_dp'))
except Exception as e: print('parse', e)
PY
