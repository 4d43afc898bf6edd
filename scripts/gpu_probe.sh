#!/bin/bash
# per-frame replay times for several simulator instances + the full GPU test suite + one bench line
O=gpurun_out
python scripts/replay_probe.py 256 4 50 > $O/probe_replay.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -15 > $O/probe_tests.txt
timeout 300 python bench.py --no-cpu-baseline --no-sweep --batched-sequences 0 > $O/probe_bench_n256.json 2> $O/probe_bench_n256.err
( export EQVIO_B200_LIB=$PWD/eqvio_b200/lib/libeqvio_b200_tl.so; EQVIO_TL_FLUSH=1 timeout 120 python scripts/timeline.py 256 2 1 > $O/probe_timeline_n256_coldL2.txt 2>&1 )
grep -v "^    " $O/probe_replay.txt; tail -3 $O/probe_tests.txt; head -c 300 $O/probe_bench_n256.json
