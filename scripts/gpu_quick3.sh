#!/bin/bash
set -u
O=gpurun_out; T=${1:-r2l}
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
for n in 256 64 1024; do
timeout 300 python bench.py --landmarks $n --no-sweep --no-cpu-baseline --batched-sequences 0 > $O/${T}_bench_n$n.json 2>/dev/null
python -c "
import json; d=json.loads(open('$O/${T}_bench_n$n.json').read().strip().splitlines()[-1]); k=d['roofline']['kernels']['prop_ll']; print('N=$n value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'prop_ll us', round(k['avg_launch_us'],2), 'frac', round(k['frac'],3))"
done
