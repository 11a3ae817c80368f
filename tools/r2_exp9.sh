#!/usr/bin/env bash
# round-2 experiment 9 (GPU box): TMA bulk-copy leaf kernels, pair GEMM tiling fix, separate reduce, PDL off
set -u
O=gpurun_out/exp9; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_full_size.py tests/test_cpp_api.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -8 > $O/tests.log
B="--no-secondary --no-cpu-baseline --steps 5 --warmup 3"
python bench.py $B --config 2 > $O/c2.json 2>> $O/err.log
python bench.py $B --config 1 > $O/c1.json 2>> $O/err.log
python bench.py $B --config 2 --shard-of 8 > $O/c2s8.json 2>> $O/err.log
python bench.py $B --config 4 --shard-of 8 > $O/c4s8.json 2>> $O/err.log
python bench.py $B --config 3 --shard-of 8 > $O/c3s8.json 2>> $O/err.log
python bench.py $B --config 3 > $O/c3.json 2>> $O/err.log
python bench.py $B --config 4 > $O/c4.json 2>> $O/err.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_c2s8.csv python tools/ncu_target_cfg.py 2 8 5 > $O/ncu_c2s8.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_cfg1.csv python tools/ncu_target_cfg.py 1 1 10 > $O/ncu_cfg1.log 2>&1
tail -3 $O/tests.log
