#!/bin/bash
# Lazy-downdate A/B at N = 1024 (run under gpurun): bit-identity check, bench lines per (M, persistent deferred launch), timelines.
set -u
O=gpurun_out; T=${1:-r2h}
timeout 300 python scripts/lazy_check.py 1024 3 1 > $O/${T}_lazy_check.txt 2>&1; echo "check rc=$?" >> $O/${T}_lazy_check.txt
tail -9 $O/${T}_lazy_check.txt
for RP in 2 1 0; do for M in 0 1 2 3 4; do
  if [ $M = 0 ] && [ $RP != 0 ]; then continue; fi
  EQVIO_B200_REST_PERSIST=$RP EQVIO_B200_LAZY=$M timeout 300 python bench.py --landmarks 1024 --no-sweep --no-cpu-baseline --batched-sequences 0 --profile-steps 0 > $O/${T}_bench_n1024_lazy${M}_rp$RP.json 2> $O/${T}_bench_n1024_lazy${M}_rp$RP.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/${T}_bench_n1024_lazy${M}_rp$RP.json").read().strip().splitlines()[-1])
    print("M=$M rp=$RP", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 4))
except Exception as e:
    print("M=$M rp=$RP failed", e)
PY
done; done
( export EQVIO_B200_LIB=$PWD/eqvio_b200/lib/libeqvio_b200_tl.so
  for M in 1 2; do EQVIO_B200_LAZY=$M timeout 120 python scripts/timeline.py 1024 2 1 > $O/${T}_timeline_n1024_lazy$M.txt 2>&1; done )
sed -n 40,58p $O/${T}_timeline_n1024_lazy1.txt
sed -n 40,62p $O/${T}_timeline_n1024_lazy2.txt
