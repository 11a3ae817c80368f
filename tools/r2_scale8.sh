#!/usr/bin/env bash
# round-2 scaling check on one 8-GPU box: the driver's invocation at N = 8 and N = 4
set -u
O=gpurun_out/scale8; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > $O/n8.json 2> $O/n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > $O/n4.json 2> $O/n4.err
tail -2 $O/n8.err; tail -2 $O/n4.err; head -c 600 $O/n8.json
