#!/usr/bin/env bash
# round-2 check on two GPUs: the multi-GPU tests, then the driver's invocation at N = 2 (own arm and reference arm)
set -u
O=gpurun_out/scale2; mkdir -p $O
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > $O/n2.json 2> $O/n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/n2_ref.json 2> $O/n2_ref.err
tail -3 $O/tests.log; tail -2 $O/n2.err; head -c 400 $O/n2.json
