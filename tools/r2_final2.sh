#!/usr/bin/env bash
# last check of the round on the final tree: the driver's GPU test command, smoke(), the default bench line
set -u
O=gpurun_out/final2; mkdir -p $O
timeout 1500 python -m pytest tests/ -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 900 python bench.py > $O/n1_default_flags.json 2> $O/n1.err
tail -2 $O/tests.log; tail -1 $O/smoke.log; head -c 250 $O/n1_default_flags.json
