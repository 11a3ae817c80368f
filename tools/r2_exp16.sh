#!/usr/bin/env bash
# round-2 experiment 16 (GPU box): narrow column tail (second instance of the contraction kernels) on / off
set -u
O=gpurun_out/exp16; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_full_size.py tests/test_cpp_api.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -8 > $O/tests.log
B="--no-secondary --no-cpu-baseline --steps 5 --warmup 3"
for v in narrow nonarrow; do
  if [ $v = nonarrow ]; then export CALS_B200_NO_NARROW=1; else unset CALS_B200_NO_NARROW; fi
  python bench.py $B --config 2 --shard-of 8 > $O/c2s8_$v.json 2>> $O/err.log
  python bench.py $B --config 1 > $O/c1_$v.json 2>> $O/err.log
  python bench.py $B --config 2 --shard-of 4 > $O/c2s4_$v.json 2>> $O/err.log
  python bench.py $B --config 3 --shard-of 8 > $O/c3s8_$v.json 2>> $O/err.log
done
unset CALS_B200_NO_NARROW
python bench.py $B --config 2 > $O/c2_narrow.json 2>> $O/err.log
python bench.py $B --config 4 --shard-of 8 > $O/c4s8_narrow.json 2>> $O/err.log
tail -4 $O/tests.log
