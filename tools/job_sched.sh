# usage: bash tools/job_sched.sh <tag> [env switches to A/B, one bench run each]
set -x
T=${1:-s}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest_$T.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_$T.log
tail -n 5 gpurun_out/r2_pytest_$T.log
timeout 600 python bench.py > gpurun_out/r2_bench_$T.json 2> gpurun_out/r2_bench_$T.err
for E in "$@"; do
  env $E timeout 600 python bench.py > gpurun_out/r2_bench_${T}_$E.json 2> gpurun_out/r2_bench_${T}_$E.err
done
timeout 600 python bench.py > gpurun_out/r2_bench_${T}b.json 2> gpurun_out/r2_bench_${T}b.err
timeout 600 python tools/step_timeline.py > gpurun_out/r2_timeline_$T.json 2> gpurun_out/r2_timeline_$T.err
timeout 900 python bench.py --workload config4 > gpurun_out/r2_bench_c4_$T.json 2> gpurun_out/r2_bench_c4_$T.err
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_*$T*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
    except Exception as e: print(f, 'parse', e)
PY
