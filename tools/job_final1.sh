set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_pytest_q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_q.log
timeout 400 python bench.py > gpurun_out/r2_bench_q.json 2> gpurun_out/r2_bench_q.err
timeout 600 python bench.py --workload config4 > gpurun_out/r2_bench_c4_q.json 2> gpurun_out/r2_bench_c4_q.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_bench_c2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_launches_ncu.log 2>&1
gzip -f gpurun_out/r2_launches_bench_c2.csv
tail -n 3 gpurun_out/r2_pytest_q.log
python - <<PY
import json
for f in ('q','c4_q'):
    try:
        d=json.loads(open('gpurun_out/r2_bench_%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
    except Exception as e: print(f,'parse',e)
PY
ls -la gpurun_out/r2_launches_bench_c2.csv.gz
