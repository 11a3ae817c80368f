#!/usr/bin/env bash
# round-2 experiment 17 (GPU box): taller MTTKRP tiles (CALS_B200_WM_MAX) on narrow shards
set -u
O=gpurun_out/exp17; mkdir -p $O
B="--no-secondary --no-cpu-baseline --steps 5 --warmup 3"
for wm in 6 7 8; do
  export CALS_B200_WM_MAX=$wm
  python bench.py $B --config 2 --shard-of 8 > $O/c2s8_wm$wm.json 2>> $O/err.log
  python bench.py $B --config 4 --shard-of 8 > $O/c4s8_wm$wm.json 2>> $O/err.log
  python bench.py $B --config 2 --shard-of 4 > $O/c2s4_wm$wm.json 2>> $O/err.log
  python bench.py $B --config 1 > $O/c1_wm$wm.json 2>> $O/err.log
done
export CALS_B200_WM_MAX=8
python bench.py $B --config 2 > $O/c2_wm8.json 2>> $O/err.log
python bench.py $B --config 3 --shard-of 8 > $O/c3s8_wm8.json 2>> $O/err.log
