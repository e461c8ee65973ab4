set -x
mkdir -p gpurun_out
P="python tools/prof_kernels.py --compact --iters 5"
for t in 0 2 4 6 1; do
  IGCN_ATTN_TPI=$t timeout 300 $P --what attn --B 512 --R 90 > gpurun_out/r2_attn_tpi$t.log 2>&1
  grep -E "bwd|cross_attn" gpurun_out/r2_attn_tpi$t.log | head -6
done
for t in 4 6; do
  IGCN_ATTN_TPI=$t timeout 600 python bench.py > gpurun_out/r2_bench_tpi$t.json 2> gpurun_out/r2_bench_tpi$t.err
done
timeout 600 python bench.py > gpurun_out/r2_bench_tpi0.json 2> gpurun_out/r2_bench_tpi0.err
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_tpi*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=[x for x in d.get('kernels_cupti',[]) if 'bwd2' in x.get('name','')]
        print(f, d['value'], d['ms_per_step'], k[:1])
    except Exception as e: print(f, 'parse', e)
PY
