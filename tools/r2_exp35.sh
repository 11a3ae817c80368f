#!/usr/bin/env bash
# tile height 7 on the wide buffers (exp17 only tried 8 there)
set -u
O=gpurun_out/exp35; mkdir -p $O
run() { local name=$1; shift; timeout 300 python bench.py --no-secondary --no-cpu-baseline --steps 5 --warmup 3 "$@" > $O/$name.json 2>> $O/err.log; }
for wm in 6 7; do
  export CALS_B200_WM_MAX=$wm
  run c2_wm$wm --config 2
  run c2s2_wm$wm --config 2 --shard-of 2
  run c4_wm$wm --config 4
  run c3_wm$wm --config 3
done
