#!/bin/bash
# Round-1 profiling pass (run under gpurun on one B200): A/B of the gather's epilogue prefetch, the ncu launch list of
# one bench step and one `--set full` capture of the gather launches of that step.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
for pf in 0 1; do
  GODE_SPMM_PREFETCH=$pf python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s4_bench_pf$pf.json 2> gpurun_out/s4_bench_pf$pf.err
  echo "prefetch=$pf rc=$?"; head -c 260 gpurun_out/s4_bench_pf$pf.json; echo
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/s4_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s4_ncu_list.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_spmm_vec -s 39 -c 13 -o gpurun_out/s4_spmm_full -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s4_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/s4_spmm_full.ncu-rep --page raw --csv > gpurun_out/s4_spmm_full_raw.csv 2>/dev/null
ls -la gpurun_out | head -30
