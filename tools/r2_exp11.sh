#!/usr/bin/env bash
# round-2 experiment 11 (GPU box): iterations per graph launch, leaf dispatch, plan cache
set -u
O=gpurun_out/exp11; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_full_size.py tests/test_cpp_api.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5 > $O/tests.log
B="--no-secondary --no-cpu-baseline --steps 5 --warmup 3"
for gb in 1 5 10; do
  export CALS_B200_GRAPH_BATCH=$gb
  python bench.py $B --config 1 > $O/c1_gb$gb.json 2>> $O/err.log
  python bench.py $B --config 2 --shard-of 8 > $O/c2s8_gb$gb.json 2>> $O/err.log
  python bench.py $B --config 2 > $O/c2_gb$gb.json 2>> $O/err.log
done
unset CALS_B200_GRAPH_BATCH
python bench.py $B --config 3 > $O/c3.json 2>> $O/err.log
python bench.py $B --config 4 --shard-of 8 > $O/c4s8.json 2>> $O/err.log
python bench.py $B --config 3 --shard-of 8 > $O/c3s8.json 2>> $O/err.log
tail -3 $O/tests.log
