#!/usr/bin/env bash
# round-2 experiment 12 (GPU box): zero-padded update kernel (unmasked DMMA loops)
set -u
O=gpurun_out/exp12; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_full_size.py tests/test_cpp_api.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -8 > $O/tests.log
export CALS_B200_UPDATE_PROF=1
for cfg in "2 8" "1 1"; do
  python tools/ncu_target_cfg.py $cfg 4 > $O/prof_$(echo $cfg | tr ' ' _).log 2>&1
done
unset CALS_B200_UPDATE_PROF
B="--no-secondary --no-cpu-baseline --steps 5 --warmup 3"
python bench.py $B --config 2 > $O/c2.json 2>> $O/err.log
python bench.py $B --config 2 --shard-of 8 > $O/c2s8.json 2>> $O/err.log
python bench.py $B --config 1 > $O/c1.json 2>> $O/err.log
python bench.py $B --config 3 > $O/c3.json 2>> $O/err.log
python bench.py $B --config 4 --shard-of 8 > $O/c4s8.json 2>> $O/err.log
python bench.py $B --config 2 --nnls > $O/c2_nnls.json 2>> $O/err.log
cat $O/prof_2_8.log $O/prof_1_1.log | grep prof; tail -4 $O/tests.log
