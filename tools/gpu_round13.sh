#!/bin/bash
# full GPU suite, bench (one graph / split graphs), fresh ncu --set full capture of the dominant GEMM launch
mkdir -p gpurun_out
rm -f gpurun_out/r13_*
timeout 1500 python -m pytest tests/ -q -m gpu > gpurun_out/r13_pytest.log 2>&1
tail -6 gpurun_out/r13_pytest.log
TFPP_GEMM_DUMP=gpurun_out/r13_gemm_shapes.txt timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r13_bench_n1.json 2> gpurun_out/r13_bench_n1.err
tail -c 1000 gpurun_out/r13_bench_n1.json; tail -3 gpurun_out/r13_bench_n1.err
TFPP_SPLIT_GRAPH=1 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r13_bench_split.json 2> gpurun_out/r13_bench_split.err
tail -c 300 gpurun_out/r13_bench_split.json; tail -3 gpurun_out/r13_bench_split.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'conv_gemm' -s 3 -c 1 -o gpurun_out/r13_qkv -f python tools/one_gemm.py qkv > gpurun_out/r13_ncu_qkv.log 2>&1
ls -la gpurun_out | grep r13_
