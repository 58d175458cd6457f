#!/bin/bash
# per-shape GEMM times of one step + ncu --set full captures of the kernels under investigation
mkdir -p gpurun_out
rm -f gpurun_out/r9_*
TFPP_GEMM_DUMP=gpurun_out/r9_gemm_shapes.txt timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r9_bench_n1.json 2> gpurun_out/r9_bench_n1.err
tail -c 900 gpurun_out/r9_bench_n1.json
for w in qkv s1stats smallc grouped; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_gemm|smallc_conv' -s 3 -c 1 -o gpurun_out/r9_$w -f python tools/one_gemm.py $w > gpurun_out/r9_ncu_$w.log 2>&1
done
ls -la gpurun_out | grep r9_
