#!/usr/bin/env bash
set -u
O=gpurun_out/exp32; mkdir -p $O
ncu --set full --import-source on --clock-control none -k 'regex:pair_gemm_kernel' -s 2 -c 1 -f -o $O/p python tools/ncu_target_cfg.py 2 1 3 > $O/p.log 2>&1
python tools/ncu_summary.py full $O/p.ncu-rep $O/pair.json > /dev/null 2>&1
python tools/ncu_stalls.py $O/p.ncu-rep $O/stalls.json 400 > $O/stalls.txt 2>&1
rm -f $O/p.ncu-rep
