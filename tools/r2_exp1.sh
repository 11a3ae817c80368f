#!/usr/bin/env bash
# round-2 experiment 1 (GPU box): ragged-octet kernels vs the previous build on full and sharded configurations,
# launch lists of the latency-bound cases
set -u
O=gpurun_out/exp1; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_full_size.py -m gpu -x -q -k "mttkrp or pair_node or config2_all or config4_all or config3_all" 2>&1 | tail -5 > $O/tests.log
B="--no-secondary --no-cpu-baseline --steps 5 --warmup 3"
for v in base new; do
  if [ $v = base ]; then export CALS_B200_LIB=$PWD/cp-cals_b200/variants/libcals_b200_base.so; else unset CALS_B200_LIB; fi
  python bench.py $B --config 2 > $O/c2_$v.json 2>> $O/err.log
  python bench.py $B --config 2 --shard-of 8 > $O/c2s8_$v.json 2>> $O/err.log
  python bench.py $B --config 4 --shard-of 8 > $O/c4s8_$v.json 2>> $O/err.log
  python bench.py $B --config 3 --shard-of 8 > $O/c3s8_$v.json 2>> $O/err.log
  python bench.py $B --config 2 --shard-of 2 > $O/c2s2_$v.json 2>> $O/err.log
done
unset CALS_B200_LIB
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_cfg1.csv python tools/ncu_target_cfg.py 1 1 10 > $O/ncu_cfg1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_c2s8.csv python tools/ncu_target_cfg.py 2 8 5 > $O/ncu_c2s8.log 2>&1
tail -3 $O/tests.log
