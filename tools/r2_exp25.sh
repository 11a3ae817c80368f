#!/usr/bin/env bash
set -u
O=gpurun_out/exp25; mkdir -p $O
ncu --set full --import-source on --clock-control none -k 'regex:mttkrp_dmma_kernel' -s 2 -c 1 -f -o $O/st python tools/ncu_target_cfg.py 2 8 3 > $O/st.log 2>&1
python tools/ncu_stalls.py $O/st.ncu-rep $O/stalls.json 60 > $O/stalls.txt 2>&1
ncu -i $O/st.ncu-rep --page raw --csv > $O/raw.csv 2>/dev/null
rm -f $O/st.ncu-rep
