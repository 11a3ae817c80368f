#!/usr/bin/env bash
# round-2 experiment 13 (GPU box): per-kernel timing through a timed graph
set -u
O=gpurun_out/exp13; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_bench_contract.py -m gpu -x -q 2>&1 | tail -5 > $O/tests.log
B="--no-secondary --no-cpu-baseline --steps 5 --warmup 3"
python bench.py $B --config 2 > $O/c2.json 2>> $O/err.log
python bench.py $B --config 2 --shard-of 8 > $O/c2s8.json 2>> $O/err.log
python bench.py $B --config 1 > $O/c1.json 2>> $O/err.log
python bench.py $B --config 4 --shard-of 8 > $O/c4s8.json 2>> $O/err.log
python bench.py $B --config 3 --shard-of 8 > $O/c3s8.json 2>> $O/err.log
tail -3 $O/tests.log
