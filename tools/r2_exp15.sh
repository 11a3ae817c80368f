#!/usr/bin/env bash
# round-2 experiment 15 (GPU box): is the failure of the reference's randomised ALS tests tied to the prefactor branch?
set -u
O=gpurun_out/exp15; mkdir -p $O
for v in off on; do
  if [ $v = on ]; then export CALS_B200_PREFACTOR=1; else unset CALS_B200_PREFACTOR; fi
  for i in 1 2 3 4 5 6; do
    cp-cals_b200/bin/ref_test_als > $O/als_${v}_$i.log 2>&1; echo "$v $i rc=$?" >> $O/summary.txt
    cp-cals_b200/bin/test_als > $O/tals_${v}_$i.log 2>&1; echo "$v $i test_als rc=$?" >> $O/summary.txt
  done
done
unset CALS_B200_PREFACTOR
cat $O/summary.txt; grep -l "FAILED\|Failure\|failed" $O/*.log | head; for f in $(grep -l "FAILED\|Failure" $O/*.log | head -2); do echo == $f; grep -B2 -A8 "Failure\|FAILED" $f | head -40; done
