#!/usr/bin/env bash
# after the revert of the column-tail experiment: full GPU suite + shards with the buffer-width-dependent tile height
set -u
O=gpurun_out/exp26; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
run() { local name=$1; shift; timeout 300 python bench.py --no-secondary --no-cpu-baseline --steps 5 --warmup 3 "$@" > $O/$name.json 2>> $O/err.log; }
run c2s8 --config 2 --shard-of 8
run c4s8 --config 4 --shard-of 8
run c2s4 --config 2 --shard-of 4
run c2s2 --config 2 --shard-of 2
run c3s8 --config 3 --shard-of 8
run c1 --config 1
run c2 --config 2
tail -3 $O/tests.log
