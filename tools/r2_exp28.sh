#!/usr/bin/env bash
set -u
O=gpurun_out/exp28; mkdir -p $O
for v in sep fused; do
  if [ $v = sep ]; then export CALS_B200_FUSED_LEAF=0; else export CALS_B200_FUSED_LEAF=1; fi
  ncu --set full --import-source on --clock-control none -k 'regex:pair_gemm_kernel' -s 2 -c 1 -f -o $O/p_$v python tools/ncu_target_cfg.py 2 1 3 > $O/p_$v.log 2>&1
  python tools/ncu_summary.py full $O/p_$v.ncu-rep $O/pair_$v.json > /dev/null 2>&1
  python tools/ncu_stalls.py $O/p_$v.ncu-rep $O/stalls_$v.json 40 > $O/stalls_$v.txt 2>&1
  rm -f $O/p_$v.ncu-rep
done
