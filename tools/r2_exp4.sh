#!/usr/bin/env bash
# round-2 experiment 4 (GPU box): phase times of the update kernel (clock64 stamps), e2e trace of config 3
set -u
O=gpurun_out/exp4; mkdir -p $O
export CALS_B200_UPDATE_PROF=1
for cfg in "2 8" "1 1" "2 1"; do
  python tools/ncu_target_cfg.py $cfg 6 > $O/prof_$(echo $cfg | tr ' ' _).log 2>&1
done
unset CALS_B200_UPDATE_PROF
python tools/make_case.py 3 /tmp/c3.case > /dev/null
CALS_B200_TRACE=1 cp-cals_b200/bin/bench_e2e /tmp/c3.case 3 2 > $O/e2e_c3.log 2>&1
B="--no-secondary --no-cpu-baseline --steps 5 --warmup 3"
python bench.py $B --config 2 > $O/c2.json 2>> $O/err.log
python bench.py $B --config 2 --shard-of 8 > $O/c2s8.json 2>> $O/err.log
python bench.py $B --config 1 > $O/c1.json 2>> $O/err.log
cat $O/prof_*.log
