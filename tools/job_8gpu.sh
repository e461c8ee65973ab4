set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --workload config5 --subjects 1000000 --steps 10 --warmup 3 > gpurun_out/r2_bench_config5_8gpu.json 2> gpurun_out/r2_bench_config5_8gpu.err
tail -n 5 gpurun_out/r2_bench_config5_8gpu.err
cat gpurun_out/r2_bench_config5_8gpu.json | cut -c1-1500
