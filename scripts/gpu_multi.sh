#!/bin/bash
# Multi-GPU check (gpurun --gpus N): the driver's launch line for N ranks, own arm and reference arm.
set -u
N=${1:-2}; O=gpurun_out; T=${2:-m$N}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 5 --warmup 2 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
head -c 400 $O/${T}_bench.json; tail -3 $O/${T}_bench.err
