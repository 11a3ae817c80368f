#!/usr/bin/env bash
# round-2 experiment 14 (GPU box): Cholesky prefactor on a forked graph branch
set -u
O=gpurun_out/exp14; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_full_size.py tests/test_cpp_api.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -8 > $O/tests.log
B="--no-secondary --no-cpu-baseline --steps 5 --warmup 3"
for v in pref nopref; do
  if [ $v = nopref ]; then export CALS_B200_NO_PREFACTOR=1; else unset CALS_B200_NO_PREFACTOR; fi
  python bench.py $B --config 2 > $O/c2_$v.json 2>> $O/err.log
  python bench.py $B --config 1 > $O/c1_$v.json 2>> $O/err.log
  python bench.py $B --config 2 --shard-of 8 > $O/c2s8_$v.json 2>> $O/err.log
done
unset CALS_B200_NO_PREFACTOR
python bench.py $B --config 3 > $O/c3_pref.json 2>> $O/err.log
python bench.py $B --config 4 --shard-of 8 > $O/c4s8_pref.json 2>> $O/err.log
python bench.py $B --config 3 --shard-of 8 > $O/c3s8_pref.json 2>> $O/err.log
tail -3 $O/tests.log
