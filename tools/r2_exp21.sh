#!/usr/bin/env bash
set -u
O=gpurun_out/exp21; mkdir -p $O
python tools/mttkrp_csweep.py > $O/sweep_tail.jsonl 2> $O/err.log
CALS_B200_NO_TAIL=1 python tools/mttkrp_csweep.py > $O/sweep_notail.jsonl 2>> $O/err.log
