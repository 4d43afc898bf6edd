#!/bin/bash
# Multi-GPU check (gpurun --gpus N): the driver's launch line for N ranks, own arm (host cores pinned per rank / not pinned) and reference arm.
set -u
N=${1:-2}; O=gpurun_out; T=${2:-m$N}
nproc > $O/${T}_host.txt; python -c "import os; print(sorted(os.sched_getaffinity(0)))" >> $O/${T}_host.txt; nvidia-smi topo -m >> $O/${T}_host.txt 2>&1; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)" >> $O/${T}_host.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-sweep > $O/${T}_bench.json 2> $O/${T}_bench.err
EQVIO_BENCH_PIN=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-sweep --batched-sequences 0 > $O/${T}_bench_nopin.json 2> $O/${T}_bench_nopin.err
python - $O/${T}_bench.json $O/${T}_bench_nopin.json <<'PY'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["e2e"]["value"], d["e2e"].get("per_rank_ms_per_step"), d["config"].get("host_cores_rank0"))
    except Exception as e: print(f, "ERR", e)
PY
