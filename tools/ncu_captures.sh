#!/bin/bash
# ncu --set full captures (one launch each) of the kernels north_star names: fusion attention fwd/bwd, pillar scatter,
# the dominant GEMM, the haloed-tile convs; + event timings printed by one_gemm.py.  Summaries go to gpurun_out/ncu_<name>.txt
mkdir -p gpurun_out
cap() {  # name regex case
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o gpurun_out/ncu_$1 python tools/one_gemm.py $3 > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/ncu_$1.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/ncu_$1.txt 2>&1
  echo "== $1"; head -30 gpurun_out/ncu_$1.txt
}
cap attn_fwd_c576 fusion_attn_fwd attn576
cap attn_bwd_c576 fusion_attn_bwd battn576
cap attn_fwd_c1512 fusion_attn_fwd attn1512
cap pillar pillar_count pillar
cap halo_conv halo_umma_conv3x3 halo
cap smallc smallc_conv3x3 smallc
cap nms nms_rotated nms
for c in attn576 battn576 attn1512 pillar halo smallc nms targets qkv; do python tools/one_gemm.py $c time; done 2>&1 | grep -i "us\b\|ms\b" | tail -20
