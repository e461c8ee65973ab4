set -x
mkdir -p gpurun_out
P="python tools/prof_kernels.py --compact --iters 1"
timeout 300 $P --what attn --B 512 --R 90 > gpurun_out/r2_attn_c2_plain.log 2>&1
timeout 300 $P --what attn --B 8192 --R 264 > gpurun_out/r2_attn_c4_plain.log 2>&1
timeout 300 $P --what gat --B 32 --R 90 > gpurun_out/r2_gat_c1_plain.log 2>&1
timeout 300 $P --what gat --B 4096 --R 264 > gpurun_out/r2_gat_big_plain.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_mma --launch-skip 2 -c 2 -f -o gpurun_out/r2_attn_c2 $P --what attn --B 512 --R 90 > gpurun_out/r2_attn_c2_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_mma --launch-skip 2 -c 2 -f -o gpurun_out/r2_attn_c4 $P --what attn --B 8192 --R 264 > gpurun_out/r2_attn_c4_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gat_layer --launch-skip 4 -c 4 -f -o gpurun_out/r2_gat_c1 $P --what gat --B 32 --R 90 > gpurun_out/r2_gat_c1_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gat_layer --launch-skip 4 -c 4 -f -o gpurun_out/r2_gat_big $P --what gat --B 4096 --R 264 > gpurun_out/r2_gat_big_ncu.log 2>&1
tail -5 gpurun_out/r2_attn_c2_plain.log gpurun_out/r2_attn_c4_plain.log gpurun_out/r2_gat_c1_plain.log gpurun_out/r2_gat_big_plain.log
ls -la gpurun_out/*.ncu-rep
