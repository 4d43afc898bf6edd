#!/bin/bash
# Round measurement on a GPU box (run under gpurun from the repo root): tests, bench lines, ncu launch list + full captures.
# Outputs under gpurun_out/final_*; summarise here with scripts/ncu_summary.py.
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/final_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.txt 2>&1
for n in 256 64 1024; do
  timeout 600 python bench.py --landmarks $n > $O/final_bench_n$n.json 2> $O/final_bench_n$n.err
done
for c in 1 2; do timeout 600 python bench.py --landmarks 256 --coord $c --no-cpu-baseline > $O/final_bench_n256_coord$c.json 2> $O/final_bench_n256_coord$c.err; done
python scripts/nonsteady_profile.py 256 60 > $O/final_nonsteady.txt 2>&1; python scripts/nonsteady_profile.py 64 60 >> $O/final_nonsteady.txt 2>&1
python scripts/host_profile.py 256 2>&1 | head -3 > $O/final_host_profile.txt
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $O/final_bench_reference_n256.json 2> $O/final_bench_reference.err
timeout 300 python bench.py --landmarks 256 --sequences-per-gpu 16 --no-cpu-baseline > $O/final_bench_n256_r16.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/final_launches_n256.csv \
  python bench.py --steps 3 --warmup 3 --profile-steps 0 --no-cpu-baseline --no-graph > $O/final_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chunk_factor_kernel|chunk_downdate_kernel|prop_ll_kernel|observer_fused_kernel" -s 40 -c 12 \
  -o $O/final_prof_n256 -f python bench.py --steps 3 --warmup 3 --profile-steps 0 --no-cpu-baseline --no-graph > $O/final_ncu_full256.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chunk_factor_kernel|chunk_downdate_kernel|prop_ll_kernel" -s 60 -c 9 \
  -o $O/final_prof_n1024 -f python bench.py --landmarks 1024 --steps 2 --warmup 3 --profile-steps 0 --no-cpu-baseline --no-graph > $O/final_ncu_full1024.log 2>&1
export EQVIO_B200_LIB=$PWD/eqvio_b200/lib/libeqvio_b200_tl.so
for n in 64 256 1024; do timeout 120 python scripts/timeline.py $n 2 1 > $O/final_timeline_n$n.txt 2>&1; done
ls -la $O | grep final
