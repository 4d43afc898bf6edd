#!/bin/bash
# Short GPU iteration (run under gpurun from the repo root): the parity tests that cover a change, one bench line, timelines.
# usage: bash scripts/gpu_quick.sh <tag> [pytest -k expression]
set -u
O=gpurun_out; T=${1:-q}; K=${2:-"sequence_matches or fused_observer or steady_path or normal or accurate or discrete or n256 or real_data or gating or correction_evaluation"}
timeout 1200 python -m pytest tests -m gpu -q -rf -k "$K" 2>&1 | tail -25 > $O/${T}_tests.txt
timeout 300 python bench.py --no-cpu-baseline --no-sweep --batched-sequences 0 > $O/${T}_bench_n256.json 2> $O/${T}_bench_n256.err
timeout 300 python bench.py --coord 2 --no-cpu-baseline --no-sweep --batched-sequences 0 > $O/${T}_bench_n256_coord2.json 2>/dev/null
timeout 300 python bench.py --landmarks 64 --no-cpu-baseline --no-sweep --batched-sequences 0 > $O/${T}_bench_n64.json 2>/dev/null
( export EQVIO_B200_LIB=$PWD/eqvio_b200/lib/libeqvio_b200_tl.so
  for n in 64 256; do timeout 120 python scripts/timeline.py $n 2 1 > $O/${T}_timeline_n$n.txt 2>&1; EQVIO_TL_FLUSH=1 timeout 120 python scripts/timeline.py $n 2 1 > $O/${T}_timeline_n${n}_coldL2.txt 2>&1; done )
tail -4 $O/${T}_tests.txt; head -c 600 $O/${T}_bench_n256.json
