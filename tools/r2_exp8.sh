#!/usr/bin/env bash
# round-2 experiment 8 (GPU box): programmatic dependent launch on/off, fused vs separate reduce at config 1, pair GEMM tiling
set -u
O=gpurun_out/exp8; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_full_size.py tests/test_cpp_api.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5 > $O/tests.log
B="--no-secondary --no-cpu-baseline --steps 5 --warmup 3"
for v in pdl nopdl; do
  if [ $v = nopdl ]; then export CALS_B200_NO_PDL=1; else unset CALS_B200_NO_PDL; fi
  python bench.py $B --config 2 > $O/c2_$v.json 2>> $O/err.log
  python bench.py $B --config 1 > $O/c1_$v.json 2>> $O/err.log
  python bench.py $B --config 2 --shard-of 8 > $O/c2s8_$v.json 2>> $O/err.log
  python bench.py $B --config 4 --shard-of 8 > $O/c4s8_$v.json 2>> $O/err.log
done
unset CALS_B200_NO_PDL
CALS_B200_NO_FUSED_REDUCE=1 python bench.py $B --config 1 > $O/c1_sepreduce.json 2>> $O/err.log
python bench.py $B --config 3 > $O/c3_pdl.json 2>> $O/err.log
python bench.py $B --config 4 > $O/c4_pdl.json 2>> $O/err.log
tail -3 $O/tests.log
