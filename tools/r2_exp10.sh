#!/usr/bin/env bash
# round-2 experiment 10 (2 GPUs): the driver's scaling invocation at N = 2, and the multi-GPU tests
set -u
O=gpurun_out/exp10; mkdir -p $O
python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5 > $O/tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/n2.json 2> $O/n2.err
python bench.py --steps 5 --warmup 3 > $O/n1.json 2> $O/n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/ref2.json 2> $O/ref2.err
tail -3 $O/tests.log; tail -3 $O/n2.err
