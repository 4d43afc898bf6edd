#!/bin/bash
# ncu --set full on the N = 1024 downdate launches, lazy M = $2 (default 2); summary CSV only (the report stays on the box)
set -u
O=gpurun_out; T=${1:-r2h}; M=${2:-2}
EQVIO_B200_LAZY=$M timeout 900 ncu --set full --clock-control none --import-source on -k regex:"chunk_downdate_kernel|chunk_factor_kernel" -s 40 -c 14 \
  -o $O/${T}_prof_n1024_lazy$M -f python bench.py --landmarks 1024 --steps 2 --warmup 3 --profile-steps 0 --no-cpu-baseline --no-graph --no-sweep --batched-sequences 0 > $O/${T}_ncu_full1024_lazy$M.log 2>&1
python scripts/ncu_summary.py full $O/${T}_prof_n1024_lazy$M.ncu-rep $O/${T}_ncu_full_n1024_lazy$M.csv > /dev/null 2>&1
rm -f $O/${T}_prof_n1024_lazy$M.ncu-rep
cat $O/${T}_ncu_full_n1024_lazy$M.csv | cut -c1-600
