#!/bin/bash
# haloed-tile group conv: tests + model parity, per-shape GEMM times of one step, ncu --set full captures
mkdir -p gpurun_out
rm -f gpurun_out/r9_*
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "gconv" > gpurun_out/r9_gconv_tests.log 2>&1
tail -5 gpurun_out/r9_gconv_tests.log
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_train_gpu.py -q -m gpu > gpurun_out/r9_model_tests.log 2>&1
tail -8 gpurun_out/r9_model_tests.log
TFPP_GEMM_DUMP=gpurun_out/r9_gemm_shapes.txt timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r9_bench_n1.json 2> gpurun_out/r9_bench_n1.err
tail -c 900 gpurun_out/r9_bench_n1.json; tail -3 gpurun_out/r9_bench_n1.err
timeout 300 python tools/bench_elem.py 2>&1 | grep -i "gconv\|smallc" > gpurun_out/r9_bench_elem.log
cat gpurun_out/r9_bench_elem.log
for w in qkv s1stats smallc gconv; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_gemm|smallc_conv|gconv3x3_kernel' -s 3 -c 1 -o gpurun_out/r9_$w -f python tools/one_gemm.py $w > gpurun_out/r9_ncu_$w.log 2>&1
done
ls -la gpurun_out | grep r9_
