# end-of-round check on one B200 (trimmed): smoke, full GPU test suite, bench default + reference arm + config4 + config3, then the ncu
# launch list of the default bench command
set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_final.log
timeout 400 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
timeout 400 python bench.py --impl reference > gpurun_out/r2_bench_final_ref.json 2> gpurun_out/r2_bench_final_ref.err
timeout 600 python bench.py --workload config4 > gpurun_out/r2_bench_final_c4.json 2> gpurun_out/r2_bench_final_c4.err
timeout 400 python bench.py --workload config3 > gpurun_out/r2_bench_final_c3.json 2> gpurun_out/r2_bench_final_c3.err
tail -n 2 gpurun_out/r2_smoke.log
tail -n 3 gpurun_out/r2_pytest_final.log
python - <<PY
import json
for f in ('final','final_ref','final_c4','final_c3'):
    try:
        d=json.loads(open('gpurun_out/r2_bench_%s.json'%f).read().strip().splitlines()[-1])
        print(f, d.get('value'), d.get('unit'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'))
    except Exception as e: print(f,'parse', e)
PY
timeout 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_ncu_final.log 2>&1
gzip -f gpurun_out/r2_launches_final.csv
ls -la gpurun_out/r2_launches_final.csv.gz
