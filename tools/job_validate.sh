# usage: bash tools/job_validate.sh <tag>   -- full GPU test suite, kernel timings, default bench
set -x
T=${1:-x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest_$T.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_$T.log
P="python tools/prof_kernels.py --compact --iters 3"
timeout 300 $P --what attn --B 512 --R 90 > gpurun_out/r2_attn_c2_$T.log 2>&1
timeout 300 $P --what attn --B 8192 --R 264 > gpurun_out/r2_attn_c4_$T.log 2>&1
timeout 300 $P --what go --B 512 > gpurun_out/r2_go_c2_$T.log 2>&1
timeout 600 python bench.py > gpurun_out/r2_bench_$T.json 2> gpurun_out/r2_bench_$T.err
tail -n 3 gpurun_out/r2_pytest_$T.log
tail -n 4 gpurun_out/r2_attn_c2_$T.log gpurun_out/r2_attn_c4_$T.log
grep -E "go_layer|skinny|bn_act" gpurun_out/r2_go_c2_$T.log
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_$T.json').read().strip().splitlines()[-1])
    print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['us_per_launch'], d['roofline']['frac'])
except Exception as e: print('bench parse', e)
PY
if [ "$2" = "ncu" ] || [ "$2" = "gat" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ --launch-skip 5 -c 5 -f -o gpurun_out/r2_attn2_c2_$T python tools/prof_kernels.py --compact --iters 1 --what attn --B 512 --R 90 > gpurun_out/r2_attn2_c2_ncu_$T.log 2>&1
fi
if [ "$2" = "gat" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gat_ --launch-skip 4 -c 4 -f -o gpurun_out/r2_gat_c1 python tools/prof_kernels.py --compact --iters 1 --what gat --B 32 --R 90 > gpurun_out/r2_gat_c1_ncu.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gat_ --launch-skip 4 -c 4 -f -o gpurun_out/r2_gat_big python tools/prof_kernels.py --compact --iters 1 --what gat --B 4096 --R 264 > gpurun_out/r2_gat_big_ncu.log 2>&1
fi
if [ "$3" = "more" ]; then
  timeout 600 python bench.py --workload config3 > gpurun_out/r2_bench_c3_$T.json 2> gpurun_out/r2_bench_c3_$T.err
  timeout 900 python bench.py --workload config4 > gpurun_out/r2_bench_c4_$T.json 2> gpurun_out/r2_bench_c4_$T.err
  python - <<PY
import json
for f in ('c3','c4'):
    try:
        d=json.loads(open('gpurun_out/r2_bench_%s_$T.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d.get('config',{}).get('description','')[-60:])
    except Exception as e: print(f, 'parse', e)
PY
fi
