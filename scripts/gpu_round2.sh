#!/bin/bash
# Round-2 measurement on a GPU box (run under gpurun from the repo root): tests, bench lines, timelines, in-kernel stamps, ncu launch
# list + full captures, sanitizer.  Outputs under gpurun_out/<tag>_* (tag = first argument, default r2g); summarise here with scripts/ncu_summary.py / ncu_traffic.py.
set -u
O=gpurun_out; T=${1:-r2g}
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -40 > $O/${T}_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.txt 2>&1
timeout 600 python bench.py > $O/${T}_bench_n256.json 2> $O/${T}_bench_n256.err
for n in 64 1024; do timeout 600 python bench.py --landmarks $n --no-sweep > $O/${T}_bench_n$n.json 2> $O/${T}_bench_n$n.err; done
EQVIO_B200_LAZY=0 EQVIO_B200_REST_AFTER_BAND=0 timeout 300 python bench.py --landmarks 1024 --no-sweep --no-cpu-baseline --batched-sequences 0 > $O/${T}_bench_n1024_round1_order.json 2>/dev/null
timeout 300 python scripts/lazy_check.py 1024 3 1 > $O/${T}_lazy_check.txt 2>&1
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $O/${T}_bench_reference_n256.json 2> $O/${T}_bench_reference.err
for c in 1 2; do timeout 300 python bench.py --coord $c --no-cpu-baseline --no-sweep --batched-sequences 0 > $O/${T}_bench_n256_coord$c.json 2>/dev/null; done
EQVIO_B200_CORRECTION=0 timeout 300 python bench.py --no-cpu-baseline --no-sweep --batched-sequences 0 > $O/${T}_bench_n256_chunks.json 2>/dev/null
timeout 300 python bench.py --no-cpu-baseline --no-sweep --batched-correction 2 > $O/${T}_bench_n256_batched_sweep.json 2>/dev/null
python scripts/host_profile.py 256 2>&1 | head -3 > $O/${T}_host_profile.txt
python scripts/nonsteady_profile.py 256 60 > $O/${T}_nonsteady.txt 2>&1
( export EQVIO_B200_LIB=$PWD/eqvio_b200/lib/libeqvio_b200_tl.so
  for n in 64 256; do timeout 120 python scripts/timeline.py $n 2 1 > $O/${T}_timeline_n$n.txt 2>&1; EQVIO_TL_FLUSH=1 timeout 120 python scripts/timeline.py $n 2 1 > $O/${T}_timeline_n${n}_coldL2.txt 2>&1; done
  timeout 120 python scripts/timeline.py 1024 2 1 > $O/${T}_timeline_n1024.txt 2>&1
  EQVIO_B200_LAZY=0 EQVIO_B200_REST_AFTER_BAND=0 timeout 120 python scripts/timeline.py 1024 2 1 > $O/${T}_timeline_n1024_round1_order.txt 2>&1
  timeout 120 python scripts/timeline.py 256 2 1 0 > $O/${T}_timeline_n256_chunks.txt 2>&1 )
EQVIO_B200_LIB=$PWD/eqvio_b200/lib/libeqvio_b200_timing.so timeout 120 python scripts/bc_timing.py 256 > $O/${T}_bc_timing.txt 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bc_probe tests/csrc/bc_probe.cu && /tmp/bc_probe > $O/${T}_bc_probe.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${T}_launches_n256.csv \
  python bench.py --steps 3 --warmup 3 --profile-steps 0 --no-cpu-baseline --no-graph --no-sweep --batched-sequences 0 > $O/${T}_ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"bc_diag_kernel|bc_trail_kernel|bc_next_kernel|bc_panel_kernel|bc_build_kernel|prop_ll_kernel|observer_fused_kernel|riccati_prep_kernel" -s 60 -c 48 \
  -o $O/${T}_prof_n256 -f python bench.py --steps 3 --warmup 3 --profile-steps 0 --no-cpu-baseline --no-graph --no-sweep --batched-sequences 0 > $O/${T}_ncu_full256.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"chunk_factor_kernel|chunk_downdate_kernel|prop_ll_kernel" -s 60 -c 12 \
  -o $O/${T}_prof_n1024 -f python bench.py --landmarks 1024 --steps 2 --warmup 3 --profile-steps 0 --no-cpu-baseline --no-graph --no-sweep --batched-sequences 0 > $O/${T}_ncu_full1024.log 2>&1
# summaries on the box (gpurun brings back at most 64 MiB): per-launch table + DRAM traffic per kernel class, then drop the big reports
python scripts/ncu_summary.py full $O/${T}_prof_n256.ncu-rep $O/${T}_ncu_full_n256.csv > /dev/null 2>&1
python scripts/ncu_summary.py full $O/${T}_prof_n1024.ncu-rep $O/${T}_ncu_full_n1024.csv > /dev/null 2>&1
python scripts/ncu_summary.py launches $O/${T}_launches_n256.csv $O/${T}_launches_n256.md > /dev/null 2>&1
EQVIO_TRAFFIC_OUT=$O/${T}_traffic.json python scripts/ncu_traffic.py 256=$O/${T}_prof_n256.ncu-rep 1024=$O/${T}_prof_n1024.ncu-rep > /dev/null 2>&1
for k in bc_diag_kernel bc_trail_kernel prop_ll_kernel observer_fused_kernel; do python scripts/ncu_lines.py $O/${T}_prof_n256.ncu-rep $k 30 > $O/${T}_lines_$k.txt 2>&1; done
rm -f $O/${T}_prof_n256.ncu-rep $O/${T}_prof_n1024.ncu-rep
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "evaluation_orders or sequence_matches or steady_path or gating" > $O/${T}_sanitizer.txt 2>&1
tail -5 $O/${T}_sanitizer.txt
ls -la $O | grep $T | wc -l
