#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02bb_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-library-baseline > gpurun_out/r02bb_ncu_list.log 2>&1; echo "ncu list rc=$?"
GODE_RK_RUNNING=0 GODE_UNIT_T=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02bb_launches_plain.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-library-baseline > gpurun_out/r02bb_ncu_list_plain.log 2>&1; echo "ncu list plain rc=$?"
