#!/usr/bin/env bash
# warp-stall sampling of the contraction kernels at the 8-way shard of config 2 and at config 2 (source-level)
set -u
O=gpurun_out/exp18; mkdir -p $O
for t in "2 8 c2s8" "2 1 c2"; do
  set -- $t
  ncu --set full --import-source on --clock-control none -k 'regex:mttkrp_dmma_kernel|pair_gemm_kernel' -s 4 -c 3 -f -o $O/st_$3 python tools/ncu_target_cfg.py $1 $2 3 > $O/st_$3.log 2>&1
  python tools/ncu_stalls.py $O/st_$3.ncu-rep $O/stalls_$3.json 40 > $O/stalls_$3.txt 2>&1
  rm -f $O/st_$3.ncu-rep
done
ls -la $O
