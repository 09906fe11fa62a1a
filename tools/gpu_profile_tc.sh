#!/bin/bash
# ncu --set full of the tcgen05 streaming kernels (transform / input grad / weight grad) on a 2 M-node graph.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_rows_ws|k_wgrad_tc" -s 24 -c 6 -o gpurun_out/s6_tc_full -f \
  python bench.py --nodes 2000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s6_ncu_tc.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/s6_tc_full.ncu-rep --page raw --csv > gpurun_out/s6_tc_full_raw.csv 2>/dev/null
ncu -i gpurun_out/s6_tc_full.ncu-rep --page details --csv > gpurun_out/s6_tc_full_details.csv 2>/dev/null
ncu -i gpurun_out/s6_tc_full.ncu-rep --page source --csv --print-source sass > gpurun_out/s6_tc_full_source.csv 2>/dev/null
ls -la gpurun_out/s6*
