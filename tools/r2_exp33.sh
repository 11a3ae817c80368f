#!/usr/bin/env bash
# fused first leaf with its default rule: full GPU suite, shards, then the default bench line
set -u
O=gpurun_out/exp33; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
run() { local name=$1; shift; timeout 300 python bench.py --no-secondary --no-cpu-baseline --steps 5 --warmup 3 "$@" > $O/$name.json 2>> $O/err.log; }
run c2s8 --config 2 --shard-of 8
run c2s4 --config 2 --shard-of 4
run c2s2 --config 2 --shard-of 2
run c1 --config 1
run c2 --config 2
timeout 900 python bench.py --steps 10 --warmup 3 > $O/default.json 2>> $O/err.log
tail -3 $O/tests.log
