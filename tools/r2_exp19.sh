#!/usr/bin/env bash
# column tail as tail items of the last n-tile: parity tests, then A/B against CALS_B200_NO_TAIL=1
set -u
O=gpurun_out/exp24; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
run() { # name, args...
  local name=$1; shift
  timeout 300 python bench.py --no-secondary --no-cpu-baseline --steps 5 --warmup 3 "$@" > $O/$name.json 2>> $O/err.log
}
for v in tail notail; do
  if [ $v = notail ]; then export CALS_B200_NO_TAIL=1; else unset CALS_B200_NO_TAIL; fi
  run c2s8_$v --config 2 --shard-of 8
  run c4s8_$v --config 4 --shard-of 8
  run c2s4_$v --config 2 --shard-of 4
  run c3s8_$v --config 3 --shard-of 8
  run c1_$v --config 1
  run c2_$v --config 2
done
unset CALS_B200_NO_TAIL
tail -5 $O/tests.log
python tools/mttkrp_csweep.py > $O/sweep_tail.jsonl 2>> $O/err.log
CALS_B200_NO_TAIL=1 python tools/mttkrp_csweep.py 256 257 263 296 525 > $O/sweep_notail.jsonl 2>> $O/err.log
