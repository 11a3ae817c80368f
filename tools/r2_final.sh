#!/usr/bin/env bash
# final check of the round: smoke(), the default bench line and the reference arm, as the driver runs them
set -u
O=gpurun_out/final; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/n1.json 2> $O/n1.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $O/n1_ref.json 2> $O/n1_ref.err
tail -2 $O/smoke.log; head -c 300 $O/n1.json; echo; head -c 300 $O/n1_ref.json
