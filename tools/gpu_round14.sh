#!/bin/bash
# 2 GPUs: forward/backward graph + eager NCCL all-reduce + optimizer graph
mkdir -p gpurun_out
rm -f gpurun_out/r14_*
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r14_n2.json 2> gpurun_out/r14_n2.err
echo "rc=$?"; tail -c 900 gpurun_out/r14_n2.json; tail -4 gpurun_out/r14_n2.err
