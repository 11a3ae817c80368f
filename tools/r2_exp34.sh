#!/usr/bin/env bash
# pair-GEMM epilogue with one wide multiply-add per T store and a predicate-free path for full tiles
set -u
O=gpurun_out/exp34; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
run() { local name=$1; shift; timeout 300 python bench.py --no-secondary --no-cpu-baseline --steps 5 --warmup 3 "$@" > $O/$name.json 2>> $O/err.log; }
run c2 --config 2
run c2s8 --config 2 --shard-of 8
run c1 --config 1
run c2b --config 2
tail -3 $O/tests.log
