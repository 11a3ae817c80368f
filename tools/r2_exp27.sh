#!/usr/bin/env bash
# first leaf fused into the pair-GEMM epilogue: parity tests, then A/B against CALS_B200_FUSED_LEAF=0
set -u
O=gpurun_out/exp31; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
run() { local name=$1; shift; timeout 300 python bench.py --no-secondary --no-cpu-baseline --steps 5 --warmup 3 "$@" > $O/$name.json 2>> $O/err.log; }
for v in fused sep; do
  if [ $v = sep ]; then export CALS_B200_FUSED_LEAF=0; else export CALS_B200_FUSED_LEAF=1; fi
  run c2_$v --config 2
  run c2s8_$v --config 2 --shard-of 8
  run c4s8_$v --config 4 --shard-of 8
  run c1_$v --config 1
done
unset CALS_B200_FUSED_LEAF
tail -3 $O/tests.log
