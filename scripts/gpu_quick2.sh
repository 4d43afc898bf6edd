#!/bin/bash
set -u
O=gpurun_out; T=${1:-r2k}
timeout 600 python -m pytest tests/test_gpu_parity_sizes.py tests/test_gpu_parity.py -m gpu -q -x -k "lazy or n1024_matches or evaluation_orders or gat" 2>&1 | tail -3
for M in 1 0; do
EQVIO_B200_LAZY=$M timeout 300 python bench.py --landmarks 1024 --no-sweep --no-cpu-baseline --batched-sequences 0 --profile-steps 0 > $O/${T}_bench_n1024_lazy$M.json 2>/dev/null
python -c "
import json; d=json.loads(open('$O/${T}_bench_n1024_lazy$M.json').read().strip().splitlines()[-1]); print('N=1024 M=$M value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
done
EQVIO_B200_CORRECTION=0 timeout 300 python bench.py --no-cpu-baseline --no-sweep --batched-sequences 0 --profile-steps 0 > $O/${T}_bench_n256_chunks.json 2>/dev/null
python -c "
import json; d=json.loads(open('$O/${T}_bench_n256_chunks.json').read().strip().splitlines()[-1]); print('N=256 chunks value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
