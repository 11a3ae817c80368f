#!/usr/bin/env bash
# round-2 experiment 3 (GPU box): register Cholesky + wider leaf kernels; ncu --set full of the update and leaf kernels
set -u
O=gpurun_out/exp3; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_full_size.py -m gpu -x -q 2>&1 | tail -5 > $O/tests.log
B="--no-secondary --no-cpu-baseline --steps 5 --warmup 3"
python bench.py $B --config 2 > $O/c2.json 2>> $O/err.log
python bench.py $B --config 1 > $O/c1.json 2>> $O/err.log
python bench.py $B --config 2 --shard-of 8 > $O/c2s8.json 2>> $O/err.log
python bench.py $B --config 4 --shard-of 8 > $O/c4s8.json 2>> $O/err.log
python bench.py $B --config 3 > $O/c3.json 2>> $O/err.log
ncu --set full --import-source on --clock-control none -k regex:model_update -s 6 -c 3 -o $O/upd python tools/ncu_target_cfg.py 2 8 5 > $O/ncu_upd.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:model_update -s 6 -c 3 -o $O/upd_c1 python tools/ncu_target_cfg.py 1 1 5 > $O/ncu_upd_c1.log 2>&1
ncu --set full --clock-control none -k regex:pair_leaf -s 2 -c 2 -o $O/leaf python tools/ncu_target_cfg.py 2 8 3 > $O/ncu_leaf.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_cfg1.csv python tools/ncu_target_cfg.py 1 1 10 > $O/ncu_cfg1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_c2s8.csv python tools/ncu_target_cfg.py 2 8 5 > $O/ncu_c2s8.log 2>&1
tail -3 $O/tests.log
