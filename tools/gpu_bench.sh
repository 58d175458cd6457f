#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/bench*.log gpurun_out/bwd_iso.log
for t in s2.b1 s2.b2 fuse1; do
  echo "=== $t" >> gpurun_out/bwd_iso.log
  timeout 600 python -m pytest tests/test_train_gpu.py -q -m gpu -k "test_component_backward_vs_oracle_autograd and $t" -x -s 2>&1 | grep -v "^$" | tail -60 >> gpurun_out/bwd_iso.log
done
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 1200 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 1500 gpurun_out/bench_n1.json
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --batch 32 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
tail -3 gpurun_out/smoke.log
