#!/bin/bash
# 2 GPUs: whole step incl. NCCL all-reduce captured in the CUDA graph vs eager launches
mkdir -p gpurun_out
rm -f gpurun_out/r12_*
export NCCL_DEBUG=WARN
timeout 420 env TFPP_GRAPH_NCCL=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r12_n2_graph.json 2> gpurun_out/r12_n2_graph.err
echo "graph rc=$?"; tail -c 700 gpurun_out/r12_n2_graph.json; tail -4 gpurun_out/r12_n2_graph.err
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r12_n2_eager.json 2> gpurun_out/r12_n2_eager.err
echo "eager rc=$?"; tail -c 700 gpurun_out/r12_n2_eager.json; tail -4 gpurun_out/r12_n2_eager.err
