#!/bin/bash
# Final round-1 pass on one B200: GPU tests, smoke, the default bench (with cpu_baseline and e2e), the reference arm,
# the per-config timings, then the ncu launch list and a 3-launch `--set full` capture of the gather kernel.
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/f_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/f_smoke.log 2>&1; tail -1 gpurun_out/f_smoke.log
timeout 600 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/f_bench.json; echo
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "ref rc=$?"; head -c 300 gpurun_out/f_bench_ref.json; echo
timeout 600 python tools/bench_configs.py > gpurun_out/f_configs.jsonl 2> gpurun_out/f_configs.err; echo "configs rc=$?"; cut -c1-200 gpurun_out/f_configs.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/f_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/f_ncu_list.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_spmm_vec -s 42 -c 3 -o gpurun_out/f_spmm_full -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/f_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/f_spmm_full.ncu-rep --page raw --csv > gpurun_out/f_spmm_full_raw.csv 2>/dev/null
rm -f gpurun_out/f_spmm_full.ncu-rep
ls -la gpurun_out | grep " f_"
