#!/bin/bash
# ncu --set full captures (one launch each) for VERDICT item 5 / item 4: the fusion GEMM on CTA pairs and on single CTAs,
# the stage-3 1x1 conv with statistics, the streaming BatchNorm-backward reduce, the BEV lift.
mkdir -p gpurun_out
cap() {  # name regex case [env]
  timeout 300 env $4 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o gpurun_out/ncu_$1 python tools/one_gemm.py $3 > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/ncu_$1.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/ncu_$1.txt 2>&1
  echo "== $1"; head -16 gpurun_out/ncu_$1.txt
}
cap gemm_mlp_pair conv_gemm_kernel mlp TFPP_GEMM_PAIR=1
cap gemm_mlp_single conv_gemm_kernel mlp TFPP_GEMM_PAIR=0
cap gemm_c576_stats conv_gemm_kernel c576 TFPP_GEMM_PAIR=1
cap bn_bwd_reduce bn_bwd_reduce_s_kernel bnbwd_i3 X=1
cap bev_lift bev_lift_kernel bevlift X=1
for c in mlp c576 bevlift; do python tools/one_gemm.py $c time; done 2>&1 | grep "us" | tail -5
rm -f gpurun_out/ncu_*.ncu-rep.tmp
